// Flash-attention forward for sm_100a, second structure: ONE 128-row Q tile per CTA with a DOUBLE-BUFFERED score tile, clusters of two
// CTAs sharing every K/V tile through TMA multicast (same contract as attn_sm100.cuh: softmax(Q K^T * scale) V, head dim 128, no mask).
//
// Why (profiles/attn_ablation_r02.json + the per-instruction samples of profiles/ncu_r02_attn.txt): in attn_sm100.cuh the score tile S_i
// and the probabilities P_i share TMEM columns, so per Q tile  S_j -> softmax_j -> P V_j -> S_{j+1}  is one serial chain of ~2630 clocks
// carrying 1024 clocks of MMA work; two Q tiles per CTA interleave two such chains and the tensor pipe ends at 2048 / 2630 = 78 %.  The
// ablations show what the chain is made of: no softmax at all 70 ms, softmax without the exponentials 80 ms, without the P store 82 ms,
// complete 92-94 ms; the TMEM read port is not involved (1.9 KB/clk/SM measured, tools/tmem_bw.cu), and the softmax warps spend 56 % of
// their samples waiting for S.  A third chain does not fit (TMEM: 2 x (128 S + 128 O) = 512 columns).  Here S_{j+1} goes to the OTHER
// score buffer and is issued BEFORE P V_j, so the tensor pipe computes the next scores while the softmax of the current tile runs:
//
//     tensor pipe :  ... | P V_{j-1} | S_{j+1} | P V_j | S_{j+2} | ...        (S_{j+2} overwrites S_j / P_j: issued after P V_j)
//     softmax     :        [ softmax_j : from "S_j done" to "P V_j issued" = 1024 MMA clocks ]
//
//   warp 0      TMA producer : Q once; K_j / V_j tiles through 3-deep rings, each CTA loads one 64-column slab and multicasts it to both
//   warp 1      MMA issuer   : S_j = Q K_j^T into score buffer j & 1;  O += P_j V_j (A = P_j from TMEM)
//   warp 2      TMEM allocator: S_a cols 0-127, S_b 128-255, O 256-383 (fp32); bf16 P_j overwrites columns [0,64) of its own score buffer
//   warps 4-11  softmax: 8 warps x 16 rows with the 16x256b TMEM shape (a row lives in one quad: row max / sum = 2 shuffles, no exchange
//               between warps), 64 scores per thread -- half the per-tile latency of one-thread-per-row.  P goes back with the 16x128b
//               shape, two 64-key halves published separately.  Lazy O rescale as in attn_sm100.cuh.
#pragma once
#include <cuda.h>

#include "attn_sm100.cuh"

namespace b200 {

constexpr int ATT5_THREADS = 384;
constexpr int ATT5_KV_STAGES = 3;
constexpr int ATT5_SMEM_BYTES = ATT_TILE_BYTES * (1 + 2 * ATT5_KV_STAGES) + 1024 + 256;
static_assert(ATT5_SMEM_BYTES <= 227 * 1024, "shared memory");

template <int POLY>
__global__ void __launch_bounds__(ATT5_THREADS, 1)
attn_s2_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    constexpr int NS = ATT5_KV_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_TILE_BYTES;
    uint8_t* sV = sK + NS * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * ATT_TILE_BYTES);
    uint64_t* q_full = bars;                 // [1]
    uint64_t* k_full = bars + 1;             // [NS]
    uint64_t* k_empty = k_full + NS;         // [NS]  released by the MMAs of both CTAs
    uint64_t* v_full = k_empty + NS;         // [NS]
    uint64_t* v_empty = v_full + NS;         // [NS]
    uint64_t* s_full = v_empty + NS;         // [2]   per score buffer: MMA -> softmax
    uint64_t* p_full = s_full + 2;           // [2][2] per score buffer, per 64-key half: softmax -> MMA (one arrival per softmax warp)
    uint64_t* pv_done = p_full + 4;          // [1]   MMA -> softmax (O complete up to tile j)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x;
    const int head = blockIdx.y;
    const int n_kv = (p.Lk + ATT_BN - 1) / ATT_BN;
    const int q_row0 = blockIdx.z * p.Lq, k_row0 = blockIdx.z * p.Lk;      // stacked sequences (see attn_sm100.cuh)

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < NS; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
        }
        for (int i = 0; i < 2; ++i) mbar_init(&s_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 8);
        mbar_init(pv_done, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();                      // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t cta_rank = cluster_ctarank();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            const int col = head * ATT_D;
            const int r0 = q_row0 + q_tile * ATT_BM;
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_2d(sQ, &tmap_q, q_full, col, r0);
            tma_load_2d(sQ + ATT_TILE_BYTES / 2, &tmap_q, q_full, col + 64, r0);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % NS;
                const uint32_t ph = (j / NS) & 1;
                // this CTA's 64-column slab of the tile, written into both CTAs (the peer sends the other slab)
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
                tma_load_2d_mcast(sK + st * ATT_TILE_BYTES + cta_rank * (ATT_TILE_BYTES / 2), &tmap_k, &k_full[st], col + (int)cta_rank * 64,
                                  k_row0 + j * ATT_BN, 0b11);
                mbar_wait(&v_empty[st], ph ^ 1);
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
            auto issue_s = [&](int j) {                     // S_j = Q K_j^T into score buffer j & 1; then K_j's stage is free in both CTAs
                const int st = j % NS;
                mbar_wait(&k_full[st], (j / NS) & 1);
                tc_fence_after();
                const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK + st * ATT_TILE_BYTES));
                #pragma unroll
                for (int kk = 0; kk < ATT_D / 16; ++kk) {
                    const uint64_t off = (uint64_t)(((kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32) >> 4);   // slab, then 32 B per K step
                    umma_bf16_ss(tmem_base + (j & 1) * 128, dQ + off, dK + off, idesc_s, kk != 0);
                }
                umma_commit(&s_full[j & 1]);
                umma_commit_mcast(&k_empty[st], 0b11);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            if (n_kv > 1) issue_s(1);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % NS, b = j & 1;
                const uint32_t pph = (j >> 1) & 1;
                mbar_wait(&v_full[st], (j / NS) & 1);
                const uint64_t dV = umma_desc_mnmajor_sw128(smem_u32(sV + st * ATT_TILE_BYTES), ATT_TILE_BYTES / 2);
                #pragma unroll
                for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                    if (kk == 0 || kk == ATT_BN / 32) {     // keys [0,64) / [64,128) of P_j are published separately
                        mbar_wait(&p_full[b * 2 + (kk != 0)], pph);
                        tc_fence_after();
                    }
                    // 16 keys = 8 packed TMEM columns of P; V: two [128 keys][64 d] slabs, 16 keys = 16 rows = 2048 B
                    umma_bf16_ts(tmem_base + 256, tmem_base + b * 128 + kk * 8, dV + (uint64_t)((kk * 2048) >> 4), idesc_o, (j | kk) != 0);
                }
                umma_commit(pv_done);
                umma_commit_mcast(&v_empty[st], 0b11);
                if (j + 2 < n_kv) issue_s(j + 2);           // overwrites S_j / P_j: tcgen05.mma executes in issue order
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================ softmax / correction / epilogue ============================
        const int q4 = warp & 3;                              // TMEM lane quarter of this warp
        const int hh = (warp - 4) >> 2;                       // which 16 lanes of the quarter
        const int lane0 = q4 * 32 + hh * 16;
        const uint32_t lane_off = (uint32_t)lane0 << 16;
        const int qd = lane & 3, rr = lane >> 2;              // quad lane -> column pair, quad index -> row
        const uint32_t tO = tmem_base + lane_off + 256;
        float m0 = -INFINITY, m1 = -INFINITY;                 // reference maxima (log2 domain) of rows lane0 + rr and lane0 + rr + 8
        float l0 = 0.f, l1 = 0.f;                             // this thread's share of the row sums
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
        for (int j = 0; j < n_kv; ++j) {
            const int b = j & 1;
            const uint32_t tS = tmem_base + lane_off + b * 128;
            mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            uint32_t v[64];                                   // v[4i], v[4i+1]: row 0, keys 8i + 2qd, +1;  v[4i+2], v[4i+3]: row 1, same keys
            tmem_ld_16x256b_x8(tS, v);
            tmem_ld_16x256b_x8(tS + 64, v + 32);
            tmem_ld_wait();
            const int valid = p.Lk - j * ATT_BN;              // >= 128 except for the last, partial tile
            if (valid < ATT_BN) {
                #pragma unroll
                for (int i = 0; i < 16; ++i) {
                    #pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (8 * i + 2 * qd + e >= valid) { v[4 * i + e] = 0xff800000u; v[4 * i + 2 + e] = 0xff800000u; }   // -inf
                }
            }
            float mx0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), mx1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
            #pragma unroll
            for (int i = 1; i < 16; ++i) {
                mx0 = fmax3(mx0, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]));
                mx1 = fmax3(mx1, __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            mx0 *= p.scale_log2; mx1 *= p.scale_log2;          // scale > 0, so max commutes with the scaling
            // ---- lazy rescale: keep the old reference max unless a row max grew by more than 8 (factor 256)
            const bool need0 = mx0 > m0 + 8.0f, need1 = mx1 > m1 + 8.0f;
            if (__any_sync(0xffffffffu, need0 || need1)) {
                const float mn0 = need0 ? mx0 : m0, mn1 = need1 ? mx1 : m1;
                const float a0 = ex2_approx(m0 - mn0), a1 = ex2_approx(m1 - mn1);     // first tile: exp2(-inf) = 0
                if (j > 0) {
                    mbar_wait(pv_done, (j - 1) & 1);           // O must be complete (P V_{j-1}) before it is rescaled in TMEM
                    tc_fence_after();
                    #pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld_16x256b_x8(tO + c * 64, o);
                        tmem_ld_wait();
                        #pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            o[4 * i] = __float_as_uint(__uint_as_float(o[4 * i]) * a0);
                            o[4 * i + 1] = __float_as_uint(__uint_as_float(o[4 * i + 1]) * a0);
                            o[4 * i + 2] = __float_as_uint(__uint_as_float(o[4 * i + 2]) * a1);
                            o[4 * i + 3] = __float_as_uint(__uint_as_float(o[4 * i + 3]) * a1);
                        }
                        tmem_st_16x256b_x8(tO + c * 64, o);
                    }
                    tmem_st_wait();
                }
                l0 *= a0; l1 *= a1;
                m0 = mn0; m1 = mn1;
            }
            // ---- P = exp2(s * scale - m) -> bf16, two keys per 32-bit column of the same score buffer, published per 64-key half
            const uint64_t nm0 = pack_f32x2(-m0, -m0), nm1 = pack_f32x2(-m1, -m1);
            uint64_t acc0 = 0ull, acc1 = 0ull;                 // packed partial sums (+0.0f bit pattern)
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t pk[16];
                #pragma unroll
                for (int ii = 0; ii < 8; ++ii) {
                    const int i = 8 * h + ii;
                    const uint64_t x0 = fma_f32x2(pack_f32x2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), sc2, nm0);
                    const uint64_t x1 = fma_f32x2(pack_f32x2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), sc2, nm1);
                    float e0, e1, e2, e3;
                    if (POLY > 0 && (2 * ii) % (POLY > 0 ? POLY : 1) == POLY - 1) {
                        ex2_poly3_x2(x0, e0, e1);
                    } else {
                        float a, c; unpack_f32x2(x0, a, c);
                        e0 = ex2_approx(a); e1 = ex2_approx(c);
                    }
                    if (POLY > 0 && (2 * ii + 1) % (POLY > 0 ? POLY : 1) == POLY - 1) {
                        ex2_poly3_x2(x1, e2, e3);
                    } else {
                        float a, c; unpack_f32x2(x1, a, c);
                        e2 = ex2_approx(a); e3 = ex2_approx(c);
                    }
                    acc0 = add_f32x2(acc0, pack_f32x2(e0, e1));
                    acc1 = add_f32x2(acc1, pack_f32x2(e2, e3));
                    pk[2 * ii] = pack_bf16x2(e0, e1);          // row 0, P column 4i + qd
                    pk[2 * ii + 1] = pack_bf16x2(e2, e3);      // row 1, same column
                }
                tmem_st_16x128b_x8(tS + h * 32, pk);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b * 2 + h]);
            }
            float s0, s1, s2, s3;
            unpack_f32x2(acc0, s0, s1); unpack_f32x2(acc1, s2, s3);
            l0 += s0 + s1; l1 += s2 + s3;
        }
        // ---- epilogue: O / l
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        mbar_wait(pv_done, (n_kv - 1) & 1);
        tc_fence_after();
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        const long long g0 = (long long)q_tile * ATT_BM + lane0 + rr, g1 = g0 + 8;            // rows inside this sequence
        __nv_bfloat16* orow0 = p.out + (q_row0 + g0) * p.ldo + head * ATT_D + 2 * qd;
        __nv_bfloat16* orow1 = p.out + (q_row0 + g1) * p.ldo + head * ATT_D + 2 * qd;
        #pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_16x256b_x8(tO + c * 64, o);
            tmem_ld_wait();
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                // exchange inside the quad so that every thread stores 8 contiguous bytes: lanes with even qd take row 0 of columns
                // 8i + 4 (qd/2) .. +3, odd qd the same columns of row 1
                const uint32_t w0 = pack_bf16x2(__uint_as_float(o[4 * i]) * inv0, __uint_as_float(o[4 * i + 1]) * inv0);
                const uint32_t w1 = pack_bf16x2(__uint_as_float(o[4 * i + 2]) * inv1, __uint_as_float(o[4 * i + 3]) * inv1);
                const uint32_t give = (qd & 1) ? w0 : w1;                       // what the neighbour (qd ^ 1) needs from me
                const uint32_t got = __shfl_xor_sync(0xffffffffu, give, 1);
                const int col = c * 64 + 8 * i;
                if ((qd & 1) == 0) {
                    if (g0 < p.Lq) *reinterpret_cast<uint2*>(orow0 + col) = make_uint2(w0, got);            // row 0: my pair, then pair of qd + 1
                } else {
                    if (g1 < p.Lq) *reinterpret_cast<uint2*>(orow1 + col - 2) = make_uint2(got, w1);        // row 1: pair of qd - 1, then mine
                }
            }
            __syncwarp();
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
