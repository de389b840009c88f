// Flash-attention forward for sm_100a: softmax(Q K^T * scale) V, head dim 128, non-causal, no mask
// (reference: shared/attention.py:208-225 sdpa_wrapper called from models/wan/modules/model.py:385 / :265).
//
// One CTA per (128-row Q tile, head); 256 threads:
//   warp 0    TMA producer : Q tile once, then K_j / V_j tiles (128 keys) through two 2-deep smem rings
//   warp 1    MMA issuer   : S_j = Q K_j^T  (tcgen05.mma M=128,N=128,K=16 x8, both operands K-major SW128)
//                            O  += P_j V_j  (A = P_j from smem, B = V_j as an MN-major SW128 operand)
//   warp 2    TMEM allocator (S double-buffered: cols 0-127 / 128-255, O: cols 256-383, all fp32)
//   warps 4-7 softmax      : thread <-> row.  tcgen05.ld the S row, online softmax in the log2 domain
//                            (ex2.approx), bf16 P written to smem in the 128B-swizzled K-major layout the
//                            MMA expects, lazy O rescale (only when the running max grows by > 2^8, FA-4
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
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * 7 + 1024 + 256;   // Q, K x2, V x2, P x2

__global__ void __launch_bounds__(256, 1)
attn_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_TILE_BYTES;            // 2 stages
    uint8_t* sV = sK + 2 * ATT_TILE_BYTES;        // 2 stages
    uint8_t* sP = sV + 2 * ATT_TILE_BYTES;        // 2 buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE_BYTES);
    uint64_t* q_full = bars;            // [1]
    uint64_t* k_full = bars + 1;        // [2]
    uint64_t* k_empty = bars + 3;       // [2]
    uint64_t* v_full = bars + 5;        // [2]
    uint64_t* v_empty = bars + 7;       // [2]
    uint64_t* s_full = bars + 9;        // [2]  MMA -> softmax
    uint64_t* p_full = bars + 11;       // [2]  softmax -> MMA (128 arrivals)
    uint64_t* pv_done = bars + 13;      // [2]  MMA -> softmax (tile j -> barrier j&1)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x;
    const int head = blockIdx.y;
    const int n_kv = (p.Lk + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&pv_done[i], 1);
        }
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
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
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
                const int st = j & 1;
                mbar_wait(&k_full[st], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t aK = smem_u32(sK + st * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32;   // slab, then 32 B per K step
                    umma_bf16_ss(tmem_base + st * 128, umma_desc_kmajor_sw128(aQ + off), umma_desc_kmajor_sw128(aK + off),
                                 idesc_s, kk != 0);
                }
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[st]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                if (j + 1 < n_kv) issue_s(j + 1);     // S buffer (j+1)&1 was fully read before p_full[j-1] arrived
                mbar_wait(&p_full[st], ph);
                mbar_wait(&v_full[st], ph);
                tc_fence_after();
                const uint32_t aP = smem_u32(sP + st * ATT_TILE_BYTES);
                const uint32_t aV = smem_u32(sV + st * ATT_TILE_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                    const uint32_t offp = (kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32;
                    // V tile: two [128 keys][64 d] slabs; MN-major B: K step of 16 keys = 16 rows = 2048 B
                    umma_bf16_ss(tmem_O, umma_desc_kmajor_sw128(aP + offp),
                                 umma_desc_mnmajor_sw128(aV + kk * 2048, ATT_TILE_BYTES / 2), idesc_o, (j | kk) != 0);
                }
                umma_commit(&v_empty[st]);
                umma_commit(&pv_done[st]);
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
            float mx = -INFINITY;
            #pragma unroll
            for (int i = 0; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
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
            // P buffer st was last read by PV_{j-2}
            if (j >= 2) mbar_wait(&pv_done[st], ((j - 2) >> 1) & 1);
            uint8_t* prow = sP + st * ATT_TILE_BYTES + row * 128;
            float lsum = 0.f;
            #pragma unroll
            for (int c = 0; c < 16; ++c) {               // 16-byte chunks of 8 keys
                float e[8];
                #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = ex2_approx(fmaf(__uint_as_float(v[c * 8 + i]), p.scale_log2, -m_used));
                    lsum += e[i];
                }
                uint4 pk;
                pk.x = pack_bf16x2(e[0], e[1]); pk.y = pack_bf16x2(e[2], e[3]);
                pk.z = pack_bf16x2(e[4], e[5]); pk.w = pack_bf16x2(e[6], e[7]);
                const int slab = c >> 3, cc = c & 7;
                *reinterpret_cast<uint4*>(prow + slab * (ATT_TILE_BYTES / 2) + ((cc ^ (row & 7)) << 4)) = pk;
            }
            l += lsum;
            fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
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
