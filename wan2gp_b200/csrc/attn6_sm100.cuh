// Flash-attention forward for sm_100a, third structure: ONE 128-row Q tile per CTA, THREE score buffers, two softmax warpgroups that take
// the K/V tiles alternately, one shared O accumulator, clusters of two CTAs sharing every K/V tile through TMA multicast
// (same contract as attn_sm100.cuh: softmax(Q K^T * scale) V, head dim 128, no mask).
//
// What the measurements say (profiles/attn_ablation_r02.json, attn_variants_r02_call9.json, the per-instruction samples of
// profiles/ncu_r02_attn.txt, tools/tmem_bw.cu):
//  * attn_sm100.cuh (two Q tiles per CTA, S_i and P_i in the same TMEM columns) is bound by the serial chain
//    S_j -> softmax_j -> P V_j -> S_{j+1} per Q tile: ~2630 clocks for 1024 clocks of MMA work, two chains per CTA -> tensor pipe 78 %.
//    One softmax warp needs ~1500 clocks per tile for its 618 instructions (88 MUFU x 8 clk alone are 704).
//  * attn5_sm100.cuh (two score buffers, 8 warps on ONE tile) hides S_{j+1} under the softmax but is slower: with a single tile in the
//    softmax at any time the TMEM read, max, exp2 (MUFU-bound) and store phases of all warps coincide, 1430 clocks per tile.
//  Needed: two tiles in the softmax at different phases (as in the first kernel) AND a window per tile longer than its latency.
//  Three score buffers give each tile 2048 MMA clocks between "S_j done" and "P V_j issued":
//
//     tensor pipe :  ... | P V_{j-3} | S_j | P V_{j-2} | S_{j+1} | P V_{j-1} | S_{j+2} | P V_j | S_{j+3} ...   (S_{j+3} overwrites S_j / P_j)
//     warpgroup A :                    [ softmax_j (even tiles) ........................ ]
//     warpgroup B :                                      [ softmax_{j+1} (odd tiles) ........................ ]
//
//  TMEM: 3 x 128 score columns + 128 O columns = 512.  ONE O accumulator for both warpgroups means one reference maximum per row for the
//  whole pass: m_ref = the row maximum of K/V tile 0, never changed -- P = exp2(s - m_ref) may exceed 1 (bf16 / fp32 hold 2^127; relative
//  precision does not depend on the magnitude), there is no O rescale at all and, after tile 0, NO ROW-MAXIMUM PASS (104 of the 618
//  instructions per warp and tile: the softmax instruction stream, not the chain latency, is what bounds this structure -- 1305 clocks
//  per tile measured with the max pass).  A row whose partial sum exceeds 2^64 or is not finite, or whose FMA-pipe exponentials saw an
//  argument above 100, sets a flag (never seen in practice); if any row of the cluster did, BOTH CTAs run two more passes over K/V:
//  one that only records the true row maxima, one with those as the reference (every P <= 1) -- the result is exact for any input.
//
//   warp 0      TMA producer K : K_j tiles through a 3-deep ring, each CTA loads one 64-column slab and multicasts it to both
//   warp 3      TMA producer V : V_j tiles through a 2-deep ring, same multicast
//   warp 1      MMA issuer     : S_j = Q K_j^T into score buffer j % 3;  O += P_j V_j (A = P_j from TMEM, over S_j's own columns)
//   warp 2      TMEM allocator
//   warps 4-7   softmax warpgroup A (even tiles), warps 8-11 warpgroup B (odd tiles): thread <-> row, packed-fp32 arithmetic, part of the
//               exponentials on the FMA pipe (as attn_sm100.cuh PACK2 / POLY), P in two 64-key halves
#pragma once
#include <cuda.h>

#include "attn_sm100.cuh"

namespace b200 {

constexpr int ATT6_THREADS = 384;
constexpr int ATT6_K_STAGES = 3, ATT6_V_STAGES = 2;
constexpr int ATT6_SMEM_BYTES = ATT_TILE_BYTES * (1 + ATT6_K_STAGES + ATT6_V_STAGES) + 1024 + 256 + 5 * 128 * 4;
static_assert(ATT6_SMEM_BYTES <= 227 * 1024, "shared memory");

__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t cluster_addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
    return v;
}

// ex2_poly3_x2 (sm100.cuh) that also tracks the largest rounded argument it saw: t = x + 1.5 * 2^23 holds round(x) in its low mantissa
// bits, so max(t) - 1.5 * 2^23 is the largest exponent added to a polynomial value -- beyond 127 the result is garbage, not +inf.
__device__ __forceinline__ void ex2_poly3_x2_guard(uint64_t x2, float& e0, float& e1, float& tmax) {
    float x0, x1;
    unpack_f32x2(x2, x0, x1);
    x2 = pack_f32x2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
    const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f), nmagic = pack_f32x2(-12582912.0f, -12582912.0f);
    const uint64_t t2 = add_f32x2(x2, magic);
    const uint64_t r2 = add_f32x2(t2, nmagic);
    const uint64_t f2 = fma_f32x2(r2, pack_f32x2(-1.0f, -1.0f), x2);
    uint64_t p2 = fma_f32x2(pack_f32x2(0.05517105758190155f, 0.05517105758190155f), f2, pack_f32x2(0.2426096349954605f, 0.2426096349954605f));
    p2 = fma_f32x2(p2, f2, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
    p2 = fma_f32x2(p2, f2, pack_f32x2(0.9999281764030457f, 0.9999281764030457f));
    float p0, p1, t0, t1;
    unpack_f32x2(p2, p0, p1);
    unpack_f32x2(t2, t0, t1);
    tmax = fmax3(tmax, t0, t1);
    e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// SPLITW: wait for the TMEM store of each 64-key half of P and publish it at once (P V_j can start on the first half) -- or store both
// halves, wait once and publish both (one blocking wait per tile: the structure is bound by the softmax instruction stream, not by latency)
template <int POLY, bool SPLITW>
__global__ void __launch_bounds__(ATT6_THREADS, 1)
attn_s3_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    constexpr int NK = ATT6_K_STAGES, NV = ATT6_V_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_TILE_BYTES;
    uint8_t* sV = sK + NK * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NV * ATT_TILE_BYTES);
    uint64_t* q_full = bars;                 // [1]
    uint64_t* k_full = bars + 1;             // [3]
    uint64_t* k_empty = k_full + NK;         // [3]  released by the MMAs of both CTAs
    uint64_t* v_full = k_empty + NK;         // [2]
    uint64_t* v_empty = v_full + NV;         // [2]
    uint64_t* s_full = v_empty + NV;         // [3]  per score buffer: MMA -> softmax
    uint64_t* p_full = s_full + 3;           // [3][2] per score buffer, per 64-key half: softmax -> MMA (one arrival per warp of the group)
    uint64_t* o_done = p_full + 6;           // [1]  MMA -> softmax: O of this pass complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);
    uint32_t* bad_flag = tmem_slot + 1;      // a row of this CTA outran its reference maximum
    float* x_mref = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [128] reference maxima of pass 0 (group A -> group B)
    float* x_rmax = x_mref + 128;            // [2][128] per group: true row maximum over its tiles
    float* x_lsum = x_rmax + 256;            // [2][128] per group: row sum over its tiles

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x;
    const int head = blockIdx.y;
    const int n_kv = (p.Lk + ATT_BN - 1) / ATT_BN;
    const int q_row0 = blockIdx.z * p.Lq, k_row0 = blockIdx.z * p.Lk;      // stacked sequences (see attn_sm100.cuh)

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < NK; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }
        for (int i = 0; i < NV; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
        for (int i = 0; i < 3; ++i) mbar_init(&s_full[i], 1);
        for (int i = 0; i < 6; ++i) mbar_init(&p_full[i], 4);
        mbar_init(o_done, 1);
        *bad_flag = 0;
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();                      // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t cta_rank = cluster_ctarank();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t peer_flag = mapa_cluster(smem_u32(bad_flag), cta_rank ^ 1u);

    // end of a pass: every thread of both CTAs sees the same flags -> did any row of the cluster outrun its reference?
    auto pass_sync = [&]() -> bool {
        tc_fence_before();
        cluster_sync_all();
        tc_fence_after();
        return (*reinterpret_cast<volatile uint32_t*>(bad_flag) | ld_shared_cluster_u32(peer_flag)) != 0u;
    };
    // mode 0: fast pass (reference = row maximum of tile 0).  Only if a row was flagged: mode 1 = record the true row maxima (P is not
    // computed, the MMAs run on whatever the score buffers hold), mode 2 = the pass again with the true maxima as the reference.

    // g = g0 + j numbers the tiles across passes: rings and barrier phases just continue.  The service warps and the softmax warps run the
    // same pass loop in two copies so that the softmax code is dominated by its own setmaxnreg (216 registers: a whole S row per thread).
    if (warp < 4) {
    setmaxnreg_dec<72>();
    int mode = 0, pass = 0, g0 = 0;
    for (;;) {
        if (warp == 0) {
            // ============================ TMA producer: Q (once), K ============================
            if (elect_one()) {
                const int col = head * ATT_D;
                if (pass == 0) {
                    const int r0 = q_row0 + q_tile * ATT_BM;
                    mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
                    tma_load_2d(sQ, &tmap_q, q_full, col, r0);
                    tma_load_2d(sQ + ATT_TILE_BYTES / 2, &tmap_q, q_full, col + 64, r0);
                }
                for (int j = 0; j < n_kv; ++j) {
                    const int g = g0 + j, st = g % NK;
                    mbar_wait(&k_empty[st], ((g / NK) & 1) ^ 1);
                    mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
                    // this CTA's 64-column slab of the tile, written into both CTAs (the peer sends the other slab)
                    tma_load_2d_mcast(sK + st * ATT_TILE_BYTES + cta_rank * (ATT_TILE_BYTES / 2), &tmap_k, &k_full[st], col + (int)cta_rank * 64,
                                      k_row0 + j * ATT_BN, 0b11);
                }
            }
            __syncwarp();
        } else if (warp == 3) {
            // ============================ TMA producer: V ============================
            if (elect_one()) {
                const int col = head * ATT_D;
                for (int j = 0; j < n_kv; ++j) {
                    const int g = g0 + j, st = g % NV;
                    mbar_wait(&v_empty[st], ((g / NV) & 1) ^ 1);
                    mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
                    tma_load_2d_mcast(sV + st * ATT_TILE_BYTES + cta_rank * (ATT_TILE_BYTES / 2), &tmap_v, &v_full[st], col + (int)cta_rank * 64,
                                      k_row0 + j * ATT_BN, 0b11);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ============================ MMA issuer ============================
            if (elect_one()) {
                constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN, /*b_mn_major=*/false);
                constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, ATT_D, /*b_mn_major=*/true);
                const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ));
                auto issue_s = [&](int g) {                     // S_g = Q K_g^T into score buffer g % 3 (= its K stage); then the stage is free
                    const int st = g % NK;
                    mbar_wait(&k_full[st], (g / NK) & 1);
                    tc_fence_after();
                    const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK + st * ATT_TILE_BYTES));
                    #pragma unroll
                    for (int kk = 0; kk < ATT_D / 16; ++kk) {
                        const uint64_t off = (uint64_t)(((kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32) >> 4);   // slab, then 32 B per K step
                        umma_bf16_ss(tmem_base + st * 128, dQ + off, dK + off, idesc_s, kk != 0);
                    }
                    umma_commit(&s_full[st]);
                    umma_commit_mcast(&k_empty[st], 0b11);
                };
                if (pass == 0) mbar_wait(q_full, 0);
                for (int t = 0; t < 3 && t < n_kv; ++t) issue_s(g0 + t);
                for (int j = 0; j < n_kv; ++j) {
                    const int g = g0 + j, vst = g % NV, buf = g % 3;
                    const uint32_t pph = (g / 3) & 1;
                    mbar_wait(&v_full[vst], (g / NV) & 1);
                    const uint64_t dV = umma_desc_mnmajor_sw128(smem_u32(sV + vst * ATT_TILE_BYTES), ATT_TILE_BYTES / 2);
                    #pragma unroll
                    for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                        if (kk == 0 || kk == ATT_BN / 32) {     // keys [0,64) / [64,128) of P_j are published separately
                            mbar_wait(&p_full[buf * 2 + (kk != 0)], pph);
                            tc_fence_after();
                        }
                        // 16 keys = 8 packed TMEM columns of P; V: two [128 keys][64 d] slabs, 16 keys = 16 rows = 2048 B
                        umma_bf16_ts(tmem_base + 384, tmem_base + buf * 128 + kk * 8, dV + (uint64_t)((kk * 2048) >> 4), idesc_o, (j | kk) != 0);
                    }
                    umma_commit_mcast(&v_empty[vst], 0b11);
                    if (j + 3 < n_kv) issue_s(g + 3);           // overwrites S_j / P_j: tcgen05.mma executes in issue order
                }
                umma_commit(o_done);
            }
            __syncwarp();
        }
        const bool flagged = pass_sync();
        if (mode == 0) { if (!flagged) break; mode = 1; } else if (mode == 1) { mode = 2; } else break;
        g0 += n_kv;
        ++pass;
        cluster_sync_all();                  // matches the softmax warps' second barrier (exchange arrays read before they are rewritten)
    }
    } else {
    setmaxnreg_inc<216>();
    const int gi = warp >= 8 ? 1 : 0;                     // softmax warpgroup: tiles j with j % 2 == gi
    const int wq = warp & 3;                              // TMEM lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    float m_ref = 0.f, l = 0.f;
    int mode = 0, pass = 0, g0 = 0;
    for (;;) {
        {
            // ============================ softmax (warpgroup gi takes tiles j % 2 == gi) ============================
            float run_max = -INFINITY, tmax = 0.f;
            l = 0.f;
            if (mode == 0 && gi == 1) {                   // group B: wait for the reference maxima of tile 0 (group A)
                named_bar_sync(2, 256);
                m_ref = x_mref[row];
            }
            for (int j = gi; j < n_kv; j += 2) {
                const int g = g0 + j, buf = g % 3;
                const uint32_t tS = tmem_base + lane_off + buf * 128;
                mbar_wait(&s_full[buf], (g / 3) & 1);
                tc_fence_after();
                uint32_t v[128];
                #pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tS + c * 32, v + c * 32);
                tmem_ld_wait();
                const int valid = p.Lk - j * ATT_BN;       // >= 128 except for the last, partial tile
                if (valid < ATT_BN) {
                    #pragma unroll
                    for (int i = 0; i < 128; ++i)
                        if (i >= valid) v[i] = 0xff800000u;   // -inf: keys beyond Lk
                }
                if (mode == 1 || (mode == 0 && j == 0)) {
                    // row maximum of this tile: the reference of the fast pass (its tile 0), or the record of the scan pass
                    float mx4[4];
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) mx4[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
                    #pragma unroll
                    for (int i = 12; i < 124; i += 2) mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                    mx4[0] = fmax3(mx4[0], __uint_as_float(v[124]), __uint_as_float(v[125]));
                    mx4[1] = fmax3(mx4[1], __uint_as_float(v[126]), __uint_as_float(v[127]));
                    const float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]) * p.scale_log2;  // scale > 0: max commutes with the scaling
                    if (mode == 0) {                      // group A, first tile: fix the reference maxima of this pass
                        m_ref = mx;
                        x_mref[row] = mx;
                        named_bar_sync(2, 256);
                    } else {                              // scan pass: hand the score buffer back untouched
                        run_max = fmaxf(run_max, mx);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(&p_full[buf * 2]); mbar_arrive(&p_full[buf * 2 + 1]); }
                        continue;
                    }
                }
                // ---- P = exp2(s * scale - m_ref) -> bf16, two keys per 32-bit column, written over S columns [0,64) of the same buffer
                const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(-m_ref, -m_ref);
                uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};          // four independent packed accumulators (+0.0f bit pattern)
                #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    #pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[h * 64 + 2 * c]), __uint_as_float(v[h * 64 + 2 * c + 1])), sc2, nm2);
                        float e0, e1;
                        if (POLY > 0 && c % (POLY > 0 ? POLY : 1) == POLY - 1) {
                            ex2_poly3_x2_guard(x2, e0, e1, tmax);
                        } else {
                            float x0, x1;
                            unpack_f32x2(x2, x0, x1);
                            e0 = ex2_approx(x0);
                            e1 = ex2_approx(x1);
                        }
                        acc2[c & 3] = add_f32x2(acc2[c & 3], pack_f32x2(e0, e1));
                        v[h * 64 + c] = pack_bf16x2(e0, e1);
                    }
                    tmem_st_32x32b_x32(tS + h * 32, v + h * 64);
                    if (SPLITW || h == 1) {
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (SPLITW) mbar_arrive(&p_full[buf * 2 + h]);
                            else { mbar_arrive(&p_full[buf * 2]); mbar_arrive(&p_full[buf * 2 + 1]); }
                        }
                    }
                }
                float a0, a1;
                unpack_f32x2(add_f32x2(add_f32x2(acc2[0], acc2[1]), add_f32x2(acc2[2], acc2[3])), a0, a1);
                l += a0 + a1;
            }
            x_rmax[gi * 128 + row] = run_max;
            x_lsum[gi * 128 + row] = l;
            // fast pass only: 2^64 bounds every P of the row, so O cannot overflow either; the comparison is false for NaN as well
            if (mode == 0 && (!(l <= 1.8e19f) || tmax > 12582912.0f + 100.0f)) *reinterpret_cast<volatile uint32_t*>(bad_flag) = 1u;
        }
        const bool flagged = pass_sync();
        if (mode == 0) {
            if (!flagged) break;
            mode = 1;
        } else if (mode == 1) {
            m_ref = fmaxf(x_rmax[row], x_rmax[128 + row]);    // the true row maxima: no score can exceed the reference now
            mode = 2;
        } else {
            break;
        }
        // the exchange arrays are rewritten at the end of the next pass; one more cluster barrier (all threads have read them) keeps that
        // simple -- this path is never taken with real data
        g0 += n_kv;
        ++pass;
        cluster_sync_all();
    }
    {
        // ---- epilogue: O / l; warpgroup gi writes columns [64 gi, 64 gi + 64)
        const float inv_l = 1.0f / (x_lsum[row] + x_lsum[128 + row]);
        mbar_wait(o_done, pass & 1);
        tc_fence_after();
        const long long grow = (long long)q_tile * ATT_BM + row;                         // row inside this sequence
        __nv_bfloat16* orow = p.out + (q_row0 + grow) * p.ldo + head * ATT_D + gi * 64;
        const uint32_t tO = tmem_base + lane_off + 384 + gi * 64;
        #pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + c * 32, o);
            tmem_ld_wait();
            if (grow < p.Lq) {
                #pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 pk;
                    pk.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
                    pk.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
                    pk.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
                    pk.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c * 32 + i) = pk;
                }
            }
            __syncwarp();
        }
    }
    }

    tc_fence_before();
    cluster_sync_all();                      // the peer may still multicast into this CTA's shared memory / barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
