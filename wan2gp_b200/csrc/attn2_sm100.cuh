// CTA-pair (cta_group::2) flash-attention forward for sm_100a: softmax(Q K^T * scale) V, head dim 128, non-causal, no mask
// (reference: shared/attention.py:208-225 sdpa_wrapper called from models/wan/modules/model.py:385 / :265).
//
// Same algorithm and per-CTA structure as attn_sm100.cuh (two 128-row Q tiles per CTA, S and O in TMEM, P written back over S in
// TMEM and consumed by a TS MMA, two softmax warpgroups in ping-pong, lazy O rescale) -- but a CLUSTER OF TWO CTAs works on 512
// query rows of one head and every tcgen05.mma is a cta_group::2 instruction with M = 256 (Q tile i of CTA 0 stacked on Q tile i of
// CTA 1):
//   * S_i = Q_i K_j^T : B = the 128-key K tile, split by keys: CTA r stages keys [64 r, 64 r + 64)  (16 KB instead of 32 KB)
//   * O_i += P_i V_j  : B = V_j as an MN-major operand, split by head-dim columns: CTA r stages d in [64 r, 64 r + 64) (16 KB)
// so each CTA loads, stores into shared memory and feeds to the tensor core HALF of every K/V tile: per 128-key step a CTA moves
// 32 KB through TMA (was 64 KB) and the MMAs read 128 KB of shared-memory operands (was 192 KB); K/V is fetched from L2 once per
// 512 query rows (was 256).  Why it matters here: under the 1 kW cap the kernel is energy-limited (SM clock 1.4-1.6 GHz of 1.965),
// and tensor pipe, MUFU and the shared-memory pipe were co-limited at ~1024 clk per 128x128 tile (ncu r01: tensor 68 %, MUFU 68 %);
// the library kernel (cuDNN SDPA) ran the same shape 13 % faster (profiles/lib_bar_r02.json).  The freed shared memory doubles the
// K/V ring depth (4 stages).
//
// Synchronisation across the pair (leader = cluster rank 0 issues every MMA):
//   q_full, k_full[], v_full[]   leader's barriers; the TMA loads of BOTH CTAs credit their bytes there (.cta_group::2 form)
//   k_empty[], v_empty[], s_full[], pv_done[]   one copy per CTA, armed by the leader's tcgen05.commit ... multicast::cluster
//   p_full[qi][half]             leader's barriers, 8 arrivals: one per softmax warp of each CTA (remote mbarrier.arrive)
#pragma once
#include <cuda.h>

#include "attn_sm100.cuh"

namespace b200 {

constexpr int ATT2_KV_STAGES = 4;
constexpr int ATT2_HALF_BYTES = ATT_TILE_BYTES / 2;            // this CTA's half of a K or V tile: 16 KB
constexpr int ATT2_SMEM_BYTES = ATT_TILE_BYTES * ATT_QTILES + 2 * ATT2_KV_STAGES * ATT2_HALF_BYTES + 1024 + 512;

// POLY: every POLY-th exp2 of a row on the FMA pipe (0 = all on MUFU).  PACK2: softmax arithmetic with packed fp32 pairs and 3-input max
// (FFMA2 / FADD2 / FMNMX3): 3 instead of 4.5 issue slots per score.
template <int POLY, bool PACK2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATT_THREADS, 1)
attn_pair_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                      // 2 tiles [128 rows][128 d] (two 64-wide slabs each)
    uint8_t* sK = sQ + ATT_QTILES * ATT_TILE_BYTES;          // stages of [64 keys][128 d] (two [64][64] slabs of 8 KB)
    uint8_t* sV = sK + ATT2_KV_STAGES * ATT2_HALF_BYTES;     // stages of [128 keys][64 d] (one slab)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT2_KV_STAGES * ATT2_HALF_BYTES);
    uint64_t* q_full = bars;                                 // [1]   leader
    uint64_t* k_full = bars + 1;                             // [ST]  leader
    uint64_t* k_empty = k_full + ATT2_KV_STAGES;             // [ST]  per CTA
    uint64_t* v_full = k_empty + ATT2_KV_STAGES;             // [ST]  leader
    uint64_t* v_empty = v_full + ATT2_KV_STAGES;             // [ST]  per CTA
    uint64_t* s_full = v_empty + ATT2_KV_STAGES;             // [2]   per CTA
    uint64_t* p_full = s_full + 2;                           // [2][2] leader
    uint64_t* pv_done = p_full + 4;                          // [2]   per CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int q_blk = blockIdx.x >> 1;                       // 512-row query block of the pair
    const int head = blockIdx.y;
    const int n_kv = (p.Lk + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < ATT2_KV_STAGES; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pv_done[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 8);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
    setmaxnreg_dec<72>();
    if (warp == 0) {
        // ============================ TMA producer (both CTAs: own Q tiles, own halves of K / V) ============================
        if (elect_one()) {
            const int col = head * ATT_D;
            const uint32_t qf = mapa_cluster(smem_u32(q_full), 0);
            if (leader) mbar_arrive_expect_tx(q_full, 2 * ATT_QTILES * ATT_TILE_BYTES);
            #pragma unroll
            for (int i = 0; i < ATT_QTILES; ++i) {
                const int r0 = ((q_blk * 2 + (int)cta_rank) * ATT_QTILES + i) * ATT_BM;
                tma_load_2d_pair(sQ + i * ATT_TILE_BYTES, &tmap_q, qf, col, r0);
                tma_load_2d_pair(sQ + i * ATT_TILE_BYTES + ATT_TILE_BYTES / 2, &tmap_q, qf, col + 64, r0);
            }
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT2_KV_STAGES;
                const uint32_t ph = (j / ATT2_KV_STAGES) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                if (leader) mbar_arrive_expect_tx(&k_full[st], 2 * ATT2_HALF_BYTES);
                const uint32_t kf = mapa_cluster(smem_u32(&k_full[st]), 0);
                const int key0 = j * ATT_BN + (int)cta_rank * (ATT_BN / 2);           // this CTA's 64 keys of the tile
                tma_load_2d_pair(sK + st * ATT2_HALF_BYTES, &tmap_k, kf, col, key0);
                tma_load_2d_pair(sK + st * ATT2_HALF_BYTES + ATT2_HALF_BYTES / 2, &tmap_k, kf, col + 64, key0);
                mbar_wait(&v_empty[st], ph ^ 1);
                if (leader) mbar_arrive_expect_tx(&v_full[st], 2 * ATT2_HALF_BYTES);
                const uint32_t vf = mapa_cluster(smem_u32(&v_full[st]), 0);
                tma_load_2d_pair(sV + st * ATT2_HALF_BYTES, &tmap_v, vf, col + (int)cta_rank * 64, j * ATT_BN);   // this CTA's 64 d columns
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer (leader CTA) ============================
        if (leader && elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(2 * ATT_BM, ATT_BN, /*b_mn_major=*/false);
            constexpr uint32_t idesc_o = umma_idesc_bf16(2 * ATT_BM, ATT_D, /*b_mn_major=*/true);
            auto issue_s = [&](int i, int j) {              // S_i = Q_i K_j^T into TMEM cols [128 i, 128 i + 128) of both CTAs
                const uint32_t aQ = smem_u32(sQ + i * ATT_TILE_BYTES);
                const uint32_t aK = smem_u32(sK + (j % ATT2_KV_STAGES) * ATT2_HALF_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_D / 16; ++kk) {
                    const uint32_t offq = (kk >> 2) * (ATT_TILE_BYTES / 2) + (kk & 3) * 32;     // 64-wide d slab, then 32 B per K step
                    const uint32_t offk = (kk >> 2) * (ATT2_HALF_BYTES / 2) + (kk & 3) * 32;    // the K slabs hold 64 rows each
                    umma_bf16_ss_pair(tmem_base + i * 128, umma_desc_kmajor_sw128(aQ + offq), umma_desc_kmajor_sw128(aK + offk),
                                      idesc_s, kk != 0);
                }
                umma_commit_pair(&s_full[i], 0b11);
            };
            auto issue_pv = [&](int i, int j) {             // O_i += P_i V_j ; P_i = bf16 in TMEM cols [128 i, 128 i + 64) of each CTA
                const uint32_t aV = smem_u32(sV + (j % ATT2_KV_STAGES) * ATT2_HALF_BYTES);
                #pragma unroll
                for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                    if (kk == 0 || kk == ATT_BN / 32) {     // keys [0,64) / [64,128) of P_i are published separately, by both CTAs
                        mbar_wait(&p_full[i * 2 + (kk != 0)], j & 1);
                        tc_fence_after();
                    }
                    umma_bf16_ts_pair(tmem_base + 256 + i * 128, tmem_base + i * 128 + kk * 8,
                                      umma_desc_mnmajor_sw128(aV + kk * 2048, ATT2_HALF_BYTES), idesc_o, (j | kk) != 0);
                }
                umma_commit_pair(&pv_done[i], 0b11);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            umma_commit_pair(&k_empty[0], 0b11);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j % ATT2_KV_STAGES;
                const uint32_t kvph = (j / ATT2_KV_STAGES) & 1;
                const bool more = j + 1 < n_kv;
                mbar_wait(&v_full[st], kvph);
                issue_pv(0, j);
                if (more) {
                    mbar_wait(&k_full[(j + 1) % ATT2_KV_STAGES], ((j + 1) / ATT2_KV_STAGES) & 1);
                    tc_fence_after();
                    issue_s(0, j + 1);
                }
                issue_pv(1, j);
                umma_commit_pair(&v_empty[st], 0b11);
                if (more) {
                    issue_s(1, j + 1);
                    umma_commit_pair(&k_empty[(j + 1) % ATT2_KV_STAGES], 0b11);
                }
            }
        }
        __syncwarp();
    }
    } else {
        // ============================ softmax / correction / epilogue (both CTAs, own rows) ============================
        setmaxnreg_inc<216>();
        const int qi = (warp - 4) >> 2;             // Q tile of this warpgroup
        const int wq = warp & 3;                    // TMEM lane quarter
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + qi * 128;
        const uint32_t tO = tmem_base + lane_off + 256 + qi * 128;
        const uint32_t pf0 = mapa_cluster(smem_u32(&p_full[qi * 2]), 0);          // the leader's barriers
        const uint32_t pf1 = mapa_cluster(smem_u32(&p_full[qi * 2 + 1]), 0);
        float m_used = -INFINITY;
        float l = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[qi], j & 1);
            tc_fence_after();
            const int valid = p.Lk - j * ATT_BN;
            uint32_t v[128];
            #pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tS + c * 32, v + c * 32);
            tmem_ld_wait();
            if (valid < ATT_BN) {
                #pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= valid) v[i] = 0xff800000u;
            }
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
                mx *= p.scale_log2;
            }
            const bool need = mx > m_used + 8.0f;
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_used;
                const float alpha = ex2_approx(m_used - m_new);
                if (j > 0) {
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
                        if (POLY > 0 && c % (POLY > 0 ? POLY : 1) == POLY - 1) {
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
                    tmem_st_32x32b_x32(tS + h * 32, v + h * 64);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(h == 0 ? pf0 : pf1);
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
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(h == 0 ? pf0 : pf1);       // one arrival per warp on the leader's barrier
            }
            }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        }
        // ---- epilogue: O / l
        mbar_wait(&pv_done[qi], (n_kv - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const long long grow = (((long long)q_blk * 2 + cta_rank) * ATT_QTILES + qi) * ATT_BM + row;
        __nv_bfloat16* orow = p.out + grow * p.ldo + head * ATT_D;
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
    cluster_sync_all();               // the peer's shared memory / TMEM / barriers stay valid until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace b200
