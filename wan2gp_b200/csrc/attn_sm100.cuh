// Flash-attention forward for sm_100a: softmax(Q K^T * scale) V, head dim 128, non-causal, no mask
// (reference: shared/attention.py:208-225 sdpa_wrapper called from models/wan/modules/model.py:385 / :265).
//
// One CTA per (256-row Q block = two 128-row Q tiles, head); 384 threads:
//   warp 0     TMA producer : both Q tiles once, then K_j / V_j tiles (128 keys) through 2-deep smem rings
//   warp 1     MMA issuer   : S_i = Q_i K_j^T  (tcgen05.mma M=128,N=128,K=16 x8, both operands K-major SW128 smem)
//                             O_i += P_i V_j   (A = P_i read from TMEM, B = V_j as an MN-major SW128 smem operand)
//                             issue order  PV0_j, S0_{j+1}, PV1_j, S1_{j+1}: while one softmax warpgroup works on
//                             its S tile the tensor pipe runs the other Q tile's MMAs (FA-4 style ping-pong)
//   warp 2     TMEM allocator: S_0 cols 0-127, S_1 cols 128-255, O_0 cols 256-383, O_1 cols 384-511 (fp32);
//                             bf16 P_i (two keys per 32-bit column) overwrites columns [0,64) of its own S_i
//   warps 4-7  softmax warpgroup of Q tile 0,  warps 8-11 softmax warpgroup of Q tile 1: thread <-> row.
//                             Online softmax in the log2 domain (ex2.approx); the whole 128-column S row lives in
//                             registers (setmaxnreg: 216 for softmax warps, 72 for the service warps) and P goes
//                             straight back to TMEM in two 64-key halves, each published to the MMA warp as soon as
//                             it is written -- no shared-memory round trip: with both operands in smem an
//                             M=128,N=128 MMA already consumes the full 128 B/clk smem bandwidth.  Lazy O rescale (only when the
//                             running max grows by > 2^8; exact because l and O always share the same reference
//                             max).  Epilogue O / l -> bf16 -> global.
// CTAs are rasterised Q-block-fastest so all CTAs resident at one time share a head and its K/V
// (2 * Lk * 256 B = 38.7 MB at L = 75 600) stays in the 126 MB L2; K/V smem tiles are shared by both Q tiles.
#pragma once
#include <cuda.h>

#include "sm100.cuh"

namespace b200 {

struct AttnParams {
    int Lq, Lk, H;
    __nv_bfloat16* out;       // [Lq, H*128], row stride ldo
    long long ldo;
    float scale_log2;         // softmax scale * log2(e)
};

constexpr int ATT_BM = 128, ATT_BN = 128, ATT_D = 128;
constexpr int ATT_QTILES = 2;                      // Q tiles per CTA
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;      // one [128][128] bf16 tile = two [128][64] slabs
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_THREADS = 384;
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (ATT_QTILES + 2 * ATT_KV_STAGES) + 1024 + 256;

// POLY: every POLY-th column of a row takes its exp2 on the FMA pipe (0 = all on MUFU).
// SPLIT: P is handed to the MMA warp in two 64-key halves so P*V starts while the second half is still being computed.
// PACK2: softmax arithmetic on packed fp32 pairs (FFMA2 / FADD2) with 3-input max (FMNMX3): 3 instead of 4.5 issue slots per score; POLY
// then counts PAIRS (every POLY-th pair of a row takes both exponentials through ex2_poly3_x2 on the FMA pipe).  MUFU.EX2 runs at 16
// results/clk/SM: the 2 x 16 384 exponentials of one K/V step need the same 2048 clocks as its four 128x128x128 MMAs, so with all of
// them on MUFU the softmax of one Q tile cannot finish inside the MMA time of the other (ncu r01: tensor 68 %, MUFU 68 %).
// MC (launched as clusters of two CTAs = two adjacent Q blocks of one head): every K/V tile is fetched from L2 ONCE for both CTAs -- CTA r
// loads the 64-column slab r of the tile and multicasts it into both shared memories; a stage is refilled only after the MMAs of BOTH
// CTAs released it (their tcgen05.commit arrives on both "empty" barriers).  All MMAs and the whole softmax <-> MMA hand-off stay
// CTA-local (unlike attn2_sm100.cuh, whose cross-CTA P hand-off cost more than the shared operands saved); only the prefetch ring is
// coupled.  Why: K/V come from L2 (hit rate 98.4 %) at 64 KB per CTA per 2048-clock step = 4.7 KB/clk over the chip, 75 % of the ~6.3 KB/clk
// the L2 can deliver.
// ABL (measurement only, results are WRONG for ABL != 0 -- tools/attn_ab.py "limiter ablation", profiles/attn_ablation_r02.json): which unit
// bounds the K/V step.  1: exponentials replaced by a move (MUFU + polynomial off); 2: only half of each S row is read from TMEM (the other
// half is derived in registers); 3: no softmax at all (S "consumed" and P "published" at once: the MMA / shared-memory / barrier ceiling of
// this issue order); 4: S is read from TMEM but nothing is computed or stored; 5: full arithmetic, P never stored to TMEM.
template <int POLY, bool SPLIT, bool PACK2 = false, bool MC = false, int ABL = 0>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                      // 2 tiles
    uint8_t* sK = sQ + ATT_QTILES * ATT_TILE_BYTES;          // ATT_KV_STAGES stages
    uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;       // ATT_KV_STAGES stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;            // [1]
    uint64_t* k_full = bars + 1;        // [2]
    uint64_t* k_empty = bars + 3;       // [2]
    uint64_t* v_full = bars + 5;        // [2]
    uint64_t* v_empty = bars + 7;       // [2]
    uint64_t* s_full = bars + 9;        // [2]  per Q tile: MMA -> softmax, one phase per KV tile
    uint64_t* p_full = bars + 11;       // [2][2]  per Q tile, per 64-key half: softmax -> MMA (128 arrivals)
    uint64_t* pv_done = bars + 15;      // [2]  per Q tile: MMA -> softmax
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x;
    const int head = blockIdx.y;
    const int n_kv = (p.Lk + ATT_BN - 1) / ATT_BN;
    // blockIdx.z = sequence of a batch of equally long sequences stacked along the rows (CFG cond / uncond branches in one launch):
    // rows [z Lq, z Lq + Lq) of q / out and [z Lk, z Lk + Lk) of k / v.  A partial last K/V tile then reads rows of the NEXT sequence
    // instead of TMA zero fill -- finite data whose scores are masked to -inf below, exactly like the zero-filled rows.
    const int q_row0 = blockIdx.z * p.Lq, k_row0 = blockIdx.z * p.Lk;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < ATT_KV_STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], MC ? 2 : 1);        // MC: released by the MMAs of both CTAs
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], MC ? 2 : 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pv_done[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], PACK2 ? 4 : 128);     // PACK2: one arrival per softmax warp (lane 0 after __syncwarp)
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    if constexpr (MC) cluster_sync_all(); else __syncthreads();                   // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t cta_rank = MC ? cluster_ctarank() : 0u;
    const uint32_t tmem_base = *tmem_slot;

    // softmax threads keep a whole 128-column S row in registers: take registers from the 4 service warps
    if (warp < 4) {
    setmaxnreg_dec<72>();
    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            const int col = head * ATT_D;
            mbar_arrive_expect_tx(q_full, ATT_QTILES * ATT_TILE_BYTES);
            #pragma unroll
            for (int i = 0; i < ATT_QTILES; ++i) {
                const int r0 = q_row0 + (q_blk * ATT_QTILES + i) * ATT_BM;
                tma_load_2d(sQ + i * ATT_TILE_BYTES, &tmap_q, q_full, col, r0);
                tma_load_2d(sQ + i * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_q, q_full, col + 64, r0);
            }
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
                if constexpr (MC) {       // this CTA's 64-column slab of the tile, written into both CTAs (the peer sends the other slab)
                    tma_load_2d_mcast(sK + st * ATT_TILE_BYTES + cta_rank * (ATT_TILE_BYTES / 2), &tmap_k, &k_full[st], col + (int)cta_rank * 64,
                                      k_row0 + j * ATT_BN, 0b11);
                } else {
                    tma_load_2d(sK + st * ATT_TILE_BYTES, &tmap_k, &k_full[st], col, k_row0 + j * ATT_BN);
                    tma_load_2d(sK + st * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_k, &k_full[st], col + 64, k_row0 + j * ATT_BN);
                }
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
                if constexpr (MC) {
                    tma_load_2d_mcast(sV + st * ATT_TILE_BYTES + cta_rank * (ATT_TILE_BYTES / 2), &tmap_v, &v_full[st], col + (int)cta_rank * 64,
                                      k_row0 + j * ATT_BN, 0b11);
                } else {
                    tma_load_2d(sV + st * ATT_TILE_BYTES, &tmap_v, &v_full[st], col, k_row0 + j * ATT_BN);
                    tma_load_2d(sV + st * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_v, &v_full[st], col + 64, k_row0 + j * ATT_BN);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (elect_one()) {
            auto release = [&](uint64_t* bar) {       // K/V stage free: in MC mode the arrival goes to both CTAs' barriers
                if constexpr (MC) umma_commit_mcast(bar, 0b11); else umma_commit(bar);
            };
            constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN, /*b_mn_major=*/false);
            constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, ATT_D, /*b_mn_major=*/true);
            auto issue_s = [&](int i, int j) {              // S_i = Q_i K_j^T into TMEM cols [128 i, 128 i + 128)
                const uint32_t aQ = smem_u32(sQ + i * ATT_TILE_BYTES);
                const uint32_t aK = smem_u32(sK + (j % ATT_KV_STAGES) * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32;   // slab, then 32 B per K step
                    umma_bf16_ss(tmem_base + i * 128, umma_desc_kmajor_sw128(aQ + off), umma_desc_kmajor_sw128(aK + off),
                                 idesc_s, kk != 0);
                }
                umma_commit(&s_full[i]);
            };
            auto issue_pv = [&](int i, int j) {             // O_i += P_i V_j ; P_i = bf16 in TMEM cols [128 i, 128 i + 64)
                const uint32_t aV = smem_u32(sV + (j % ATT_KV_STAGES) * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                    if (kk == 0 || kk == ATT_BN / 32) {     // keys [0,64) / [64,128) of P_i are published separately
                        mbar_wait(&p_full[i * 2 + (kk != 0)], j & 1);
                        tc_fence_after();
                    }
                    // 16 keys = 8 packed TMEM columns of P; V: two [128 keys][64 d] slabs, 16 keys = 16 rows = 2048 B
                    umma_bf16_ts(tmem_base + 256 + i * 128, tmem_base + i * 128 + kk * 8,
                                 umma_desc_mnmajor_sw128(aV + kk * 2048, ATT_TILE_BYTES / 2), idesc_o, (j | kk) != 0);
                }
                umma_commit(&pv_done[i]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            release(&k_empty[0]);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t kvph = (j / ATT_KV_STAGES) & 1;
                const bool more = j + 1 < n_kv;
                // ---- Q tile 0: PV0_j then S0_{j+1} (S0_{j+1} overwrites P0_j: tcgen05.mma executes in issue order)
                mbar_wait(&v_full[st], kvph);
                issue_pv(0, j);
                if (more) {
                    mbar_wait(&k_full[(j + 1) % ATT_KV_STAGES], ((j + 1) / ATT_KV_STAGES) & 1);
                    tc_fence_after();
                    issue_s(0, j + 1);
                }
                // ---- Q tile 1
                issue_pv(1, j);
                release(&v_empty[st]);
                if (more) {
                    issue_s(1, j + 1);
                    release(&k_empty[(j + 1) % ATT_KV_STAGES]);
                }
            }
        }
        __syncwarp();
    }
    } else {
        // ============================ softmax / correction / epilogue ============================
        setmaxnreg_inc<216>();
        const int qi = (warp - 4) >> 2;             // Q tile of this warpgroup
        const int wq = warp & 3;                    // TMEM lane quarter
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + qi * 128;
        const uint32_t tO = tmem_base + lane_off + 256 + qi * 128;
        float m_used = -INFINITY;      // reference max (log2 domain) shared by l and O
        float l = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[qi], j & 1);
            tc_fence_after();
            const int valid = p.Lk - j * ATT_BN;       // >= 128 except for the last, partial tile
            uint32_t v[128];
            if constexpr (ABL == 3) {                     // MMA-only ceiling: hand the (garbage) S straight back as P
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&p_full[qi * 2]); mbar_arrive(&p_full[qi * 2 + 1]); }
                continue;
            }
            if constexpr (ABL == 2) {
                #pragma unroll
                for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(tS + c * 32, v + c * 32);
                tmem_ld_wait();
                #pragma unroll
                for (int i = 0; i < 64; ++i) v[64 + i] = v[i] ^ (uint32_t)((j & 1) << 12);      // distinct values: nothing is common-subexpression'd away
            } else {
            #pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tS + c * 32, v + c * 32);
            tmem_ld_wait();
            }
            if constexpr (ABL == 4) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&p_full[qi * 2]); mbar_arrive(&p_full[qi * 2 + 1]); }
                continue;
            }
            if (valid < ATT_BN) {
                #pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= valid) v[i] = 0xff800000u;   // -inf: keys beyond Lk (TMA zero-filled rows)
            }
            // row max with 8 independent partial maxima (ILP: at most two softmax warps share a scheduler)
            float mx;
            if constexpr (PACK2) {
                float mx4[4];
                #pragma unroll
                for (int i = 0; i < 4; ++i) mx4[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
                #pragma unroll
                for (int i = 12; i < 124; i += 2) mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                mx4[0] = fmax3(mx4[0], __uint_as_float(v[124]), __uint_as_float(v[125]));
                mx4[1] = fmax3(mx4[1], __uint_as_float(v[126]), __uint_as_float(v[127]));
                mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]) * p.scale_log2;
            } else {
                float mx8[8];
                #pragma unroll
                for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(v[i]);
                #pragma unroll
                for (int i = 8; i < 128; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(v[i]));
                mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
                mx *= p.scale_log2;                    // scale > 0, so max commutes with the scaling
            }
            // ---- lazy rescale: keep the old reference max unless the row max grew by more than 8 (factor 256)
            const bool need = mx > m_used + 8.0f;
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_used;
                const float alpha = ex2_approx(m_used - m_new);   // first tile: exp2(-inf) = 0
                if (j > 0) {
                    // O must be complete (PV_{j-1}) before it is rescaled in TMEM
                    mbar_wait(&pv_done[qi], (j - 1) & 1);
                    tc_fence_after();
                    #pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(tO + c * 32, o);
                        tmem_ld_wait();
                        #pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32b_x32(tO + c * 32, o);
                    }
                    tmem_st_wait();
                }
                l *= alpha;
                m_used = m_new;
            }
            // ---- P = exp2(s * scale - m) -> bf16, two keys per 32-bit column, written over S columns [0,64); packed in
            //      place (v[h*64 + c] <- keys h*64 + 2c, 2c+1) and published per 64-key half
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
            const float neg_m = -m_used;
            if constexpr (PACK2) {
                const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(neg_m, neg_m);
                uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};          // four independent packed accumulators (+0.0f bit pattern)
                #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    #pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[h * 64 + 2 * c]), __uint_as_float(v[h * 64 + 2 * c + 1])), sc2, nm2);
                        float e0, e1;
                        if constexpr (ABL == 1) {
                            unpack_f32x2(x2, e0, e1);
                        } else if (POLY > 0 && c % (POLY > 0 ? POLY : 1) == POLY - 1) {
                            ex2_poly3_x2(x2, e0, e1);
                        } else {
                            float x0, x1;
                            unpack_f32x2(x2, x0, x1);
                            e0 = ex2_approx(x0);
                            e1 = ex2_approx(x1);
                        }
                        acc2[c & 3] = add_f32x2(acc2[c & 3], pack_f32x2(e0, e1));
                        v[h * 64 + c] = pack_bf16x2(e0, e1);
                    }
                    if constexpr (ABL != 5) tmem_st_32x32b_x32(tS + h * 32, v + h * 64);
                    if (SPLIT || h == 1) {
                        if constexpr (ABL != 5) tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {                                  // 4 instead of 128 shared-memory barrier updates per hand-off
                            if (SPLIT) mbar_arrive(&p_full[qi * 2 + h]);
                            else { mbar_arrive(&p_full[qi * 2]); mbar_arrive(&p_full[qi * 2 + 1]); }
                        }
                    }
                }
                float a0, a1;
                unpack_f32x2(add_f32x2(add_f32x2(acc2[0], acc2[1]), add_f32x2(acc2[2], acc2[3])), a0, a1);
                ls[0] = a0 + a1;
            } else {
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                #pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float x0 = fmaf(__uint_as_float(v[h * 64 + 2 * c]), p.scale_log2, neg_m);
                    const float x1 = fmaf(__uint_as_float(v[h * 64 + 2 * c + 1]), p.scale_log2, neg_m);
                    const float e0 = (POLY > 0 && (2 * c) % (POLY > 0 ? POLY : 1) == POLY - 1) ? ex2_poly3(x0) : ex2_approx(x0);
                    const float e1 = (POLY > 0 && (2 * c + 1) % (POLY > 0 ? POLY : 1) == POLY - 1) ? ex2_poly3(x1) : ex2_approx(x1);
                    ls[c & 3] += e0 + e1;
                    v[h * 64 + c] = pack_bf16x2(e0, e1);
                }
                tmem_st_32x32b_x32(tS + h * 32, v + h * 64);
                if (SPLIT || h == 1) {
                    tmem_st_wait();
                    tc_fence_before();
                    if (SPLIT) mbar_arrive(&p_full[qi * 2 + h]);
                    else { mbar_arrive(&p_full[qi * 2]); mbar_arrive(&p_full[qi * 2 + 1]); }
                }
            }
            }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        }
        // ---- epilogue: O / l
        mbar_wait(&pv_done[qi], (n_kv - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const long long grow = ((long long)q_blk * ATT_QTILES + qi) * ATT_BM + row;          // row inside this sequence
        __nv_bfloat16* orow = p.out + (q_row0 + grow) * p.ldo + head * ATT_D;
        #pragma unroll 1
        for (int c = 0; c < 4; ++c) {
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

    tc_fence_before();
    if constexpr (MC) cluster_sync_all(); else __syncthreads();       // the peer may still multicast into this CTA's shared memory / barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
