// Flash-attention forward for sm_100a: softmax(Q K^T * scale) V, head dim 128, non-causal, no mask
// (reference: shared/attention.py:208-225 sdpa_wrapper called from models/wan/modules/model.py:385 / :265).
//
// One CTA per (128-row Q tile, head); 256 threads:
//   warp 0    TMA producer : Q tile once, then K_j / V_j tiles (128 keys) through two 2-deep smem rings
//   warp 1    MMA issuer   : S_j = Q K_j^T  (tcgen05.mma M=128,N=128,K=16 x8, both operands K-major SW128)
//                            O  += P_j V_j  (A = P_j read from TMEM, B = V_j as an MN-major SW128 operand)
//   warp 2    TMEM allocator (S double-buffered: cols 0-127 / 128-255, O: cols 256-383, all fp32; the bf16 P_j
//                            overwrites the first 64 columns of its own S_j buffer)
//   warps 4-7 softmax      : thread <-> row.  tcgen05.ld the S row, online softmax in the log2 domain
//                            (ex2.approx), bf16 P packed 2/column and tcgen05.st back to TMEM (no shared-memory
//                            round trip: with both MMA operands in smem an M=128,N=128 MMA already consumes the
//                            full 128 B/clk smem bandwidth, so P-through-smem capped the tensor pipe at ~55%),
//                            lazy O rescale (only when the running max grows by > 2^8, FA-4
//                            style; exact because l and O always share the same reference max), epilogue
//                            O / l -> bf16 -> global.
// S_{j+1} is issued before the softmax of tile j finishes, so tensor pipe and MUFU/FMA pipes overlap.
// CTAs are rasterised Q-tile-fastest so all CTAs resident at one time share a head and its K/V
// (2 * Lk * 256 B = 38.7 MB at L = 75 600) stays in the 126 MB L2.
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
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;     // one [128][128] bf16 tile = two [128][64] slabs
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES) + 1024 + 256;   // Q, K x3, V x3

__global__ void __launch_bounds__(256, 1)
attn_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_TILE_BYTES;                       // ATT_KV_STAGES stages
    uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;       // ATT_KV_STAGES stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * ATT_TILE_BYTES);
    uint64_t* q_full = bars;            // [1]
    uint64_t* k_full = bars + 1;        // [3]
    uint64_t* k_empty = bars + 4;       // [3]
    uint64_t* v_full = bars + 7;        // [3]
    uint64_t* v_empty = bars + 10;      // [3]
    uint64_t* s_full = bars + 13;       // [2]  MMA -> softmax
    uint64_t* p_full = bars + 15;       // [2]  softmax -> MMA (128 arrivals)
    uint64_t* pv_done = bars + 17;      // [2]  MMA -> softmax (tile j -> barrier j&1)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x;
    const int head = blockIdx.y;
    const int n_kv = (p.Lk + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < ATT_KV_STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&pv_done[i], 1); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_O = tmem_base + 256;

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            const int col = head * ATT_D;
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_2d(sQ, &tmap_q, q_full, col, q_blk * ATT_BM);
            tma_load_2d(sQ + ATT_TILE_BYTES / 2, &tmap_q, q_full, col + 64, q_blk * ATT_BM);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
                tma_load_2d(sK + st * ATT_TILE_BYTES, &tmap_k, &k_full[st], col, j * ATT_BN);
                tma_load_2d(sK + st * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_k, &k_full[st], col + 64, j * ATT_BN);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
                tma_load_2d(sV + st * ATT_TILE_BYTES, &tmap_v, &v_full[st], col, j * ATT_BN);
                tma_load_2d(sV + st * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_v, &v_full[st], col + 64, j * ATT_BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN, /*b_mn_major=*/false);
            constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, ATT_D, /*b_mn_major=*/true);
            const uint32_t aQ = smem_u32(sQ);
            auto issue_s = [&](int j) {
                const int st = j % ATT_KV_STAGES, sb = j & 1;
                mbar_wait(&k_full[st], (j / ATT_KV_STAGES) & 1);
                tc_fence_after();
                const uint32_t aK = smem_u32(sK + st * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32;   // slab, then 32 B per K step
                    umma_bf16_ss(tmem_base + sb * 128, umma_desc_kmajor_sw128(aQ + off), umma_desc_kmajor_sw128(aK + off),
                                 idesc_s, kk != 0);
                }
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[sb]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT_KV_STAGES, sb = j & 1;
                // S buffer (j+1)&1 still holds P_{j-1} in its first 64 columns; PV_{j-1} was issued earlier and
                // tcgen05.mma executes in issue order, so it has consumed P_{j-1} before S_{j+1} overwrites it.
                if (j + 1 < n_kv) issue_s(j + 1);
                mbar_wait(&p_full[sb], (j >> 1) & 1);
                mbar_wait(&v_full[st], (j / ATT_KV_STAGES) & 1);
                tc_fence_after();
                const uint32_t aV = smem_u32(sV + st * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                    // A = P_j in TMEM: 16 keys = 8 packed columns per K step.
                    // V tile: two [128 keys][64 d] slabs; MN-major B: K step of 16 keys = 16 rows = 2048 B
                    umma_bf16_ts(tmem_O, tmem_base + sb * 128 + kk * 8,
                                 umma_desc_mnmajor_sw128(aV + kk * 2048, ATT_TILE_BYTES / 2), idesc_o, (j | kk) != 0);
                }
                umma_commit(&v_empty[st]);
                umma_commit(&pv_done[sb]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================ softmax / correction / epilogue ============================
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        float m_used = -INFINITY;      // reference max (log2 domain) shared by l and O
        float l = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            const int st = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(&s_full[st], ph);
            tc_fence_after();
            uint32_t v[128];
            #pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tmem_base + lane_off + st * 128 + c * 32, v + c * 32);
            tmem_ld_wait();
            const int valid = p.Lk - j * ATT_BN;       // >= 128 except for the last, partial tile
            if (valid < ATT_BN) {
                #pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= valid) v[i] = 0xff800000u;   // -inf: keys beyond Lk (TMA zero-filled rows)
            }
            // 8 independent partial maxima: this warp is the only softmax warp on its scheduler, so ILP (not TLP)
            // has to hide the 4-cycle ALU latency
            float mx8[8];
            #pragma unroll
            for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(v[i]);
            #pragma unroll
            for (int i = 8; i < 128; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(v[i]));
            float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
            mx *= p.scale_log2;                        // scale > 0, so max commutes with the scaling
            // lazy rescale: keep the old reference max unless the row max grew by more than 8 (factor 256)
            const bool need = mx > m_used + 8.0f;
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_used;
                const float alpha = ex2_approx(m_used - m_new);   // first tile: exp2(-inf) = 0
                if (j > 0) {
                    // O must be complete (PV_{j-1}) before it is rescaled in TMEM
                    mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                    tc_fence_after();
                    #pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, o);
                        tmem_ld_wait();
                        #pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32b_x32(tmem_O + lane_off + c * 32, o);
                    }
                    tmem_st_wait();
                }
                l *= alpha;
                m_used = m_new;
            }
            // P_j (bf16, two keys per 32-bit column) overwrites columns [0,64) of this thread's own S_j row, which is
            // already in registers; packed in place: v[c] <- (p[2c], p[2c+1]).
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
            const float neg_m = -m_used;
            #pragma unroll
            for (int c = 0; c < 64; ++c) {
                const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * c]), p.scale_log2, neg_m));
                const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * c + 1]), p.scale_log2, neg_m));
                ls[c & 3] += e0 + e1;
                v[c] = pack_bf16x2(e0, e1);
            }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
            tmem_st_32x32b_x32(tmem_base + lane_off + st * 128, v);
            tmem_st_32x32b_x32(tmem_base + lane_off + st * 128 + 32, v + 32);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[st]);
        }
        // ---- epilogue: O / l
        const int jl = n_kv - 1;
        mbar_wait(&pv_done[jl & 1], (jl >> 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const long long grow = (long long)q_blk * ATT_BM + row;
        __nv_bfloat16* orow = p.out + grow * p.ldo + head * ATT_D;
        #pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_O + lane_off + c * 32, o);
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
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
