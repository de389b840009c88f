// CTA-pair (cta_group::2) persistent tcgen05 GEMM for sm_100a: the large linear layers of the DiT blocks.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T + bias )         bf16 operands, fp32 accumulate in TMEM
//
// A cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  CTA r of the pair stages rows [128 r, 128 r + 128) of the A
// tile and rows [128 r, 128 r + 128) of the B (weight) tile per 64-wide K step -- 32 KB per CTA and stage instead of the 48 KB of
// the single-CTA 128 x 256 kernel (gemm_sm100.cuh) -- and ONE tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) issued by the
// leader CTA reads both shared memories: every weight byte is fetched from shared memory once per 256 output rows instead of once
// per 128, which takes the operand traffic of the MMA from 96 B/clk to 64 B/clk per SM (the shared-memory pipe is shared with the
// TMA writes that refill the ring).  Each CTA receives its own 128 accumulator rows in its own TMEM (2 x 256 columns, double
// buffered so the epilogue of tile i overlaps the main loop of tile i+1).
//
// Per CTA, 256 threads:
//   warp 0   TMA producer   waits on its OWN "stage empty" barrier, issues its A / B-half loads; all transaction bytes of both
//                           CTAs are credited to the LEADER's "stage full" barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1   MMA issuer     leader CTA only; tcgen05.commit multicasts "stage empty" / "accumulator full" to both CTAs
//   warp 2   TMEM allocator tcgen05.alloc.cta_group::2 (the same warp in both CTAs)
//   warps 4-7 epilogue      tcgen05.ld of the CTA's own accumulator rows, bias / GELU / gate / residual fused, 16-byte stores
//                           (or TMA reduce-add for the fp32 residual stream); "accumulator empty" arrives on the leader's barrier
//                           from both CTAs (remote mbarrier.arrive through the cluster address of the leader's barrier)
// Reference: the reference runs these layers through cuBLASLt (models/wan/modules/model.py:322, 337, 405, 686-711).
#pragma once
#include <cuda.h>

#include "gemm_sm100.cuh"

namespace b200 {

constexpr int GEMM2_BN = 256;          // N tile of the pair (128 weight rows staged per CTA)
constexpr int GEMM2_BM = 256;          // M tile of the pair (128 rows per CTA)

template <bool EPI_TMA>
struct Gemm2Smem {
    static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;          // 16 KB: this CTA's 128 A rows x 64 K
    static constexpr int kBBytes = (GEMM2_BN / 2) * GEMM_BK * 2;   // 16 KB: this CTA's 128 weight rows x 64 K
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBytes = EPI_TMA ? 2 * GEMM_BM * 128 : 0;
    static constexpr int kStages = (204 * 1024 - kEpiBytes) / kStageBytes;     // 6 (5 with the reduce-add staging tiles)
    static constexpr int kBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <bool EPI_TMA>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
gemm_pair_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
    using S = Gemm2Smem<EPI_TMA>;
    constexpr int kStages = S::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + kStages * S::kStageBytes;          // [2][128 rows][128 B], EPI_TMA only
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + S::kEpiBytes);
    uint64_t* full_bar = bars;                    // [kStages]  used in the leader CTA only
    uint64_t* empty_bar = bars + kStages;         // [kStages]  one per CTA, armed by the leader's multicast commit
    uint64_t* tfull_bar = bars + 2 * kStages;     // [2]        one per CTA, multicast commit
    uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]      leader only: 8 arrivals (4 epilogue warps x 2 CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();               // barriers of BOTH CTAs are initialised before any remote arrive / complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.m_tiles * p.n_tiles;          // pair tiles (256 x 256)
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        const int per_group = p.m_tiles * p.n_group;
        const int g = tile / per_group;
        const int r = tile - g * per_group;
        const int gw = min(p.n_group, p.n_tiles - g * p.n_group);
        m_blk = r / gw;
        n_blk = g * p.n_group + (r - m_blk * gw);
    };

    if (warp == 0) {
        // ============================ TMA producer (both CTAs) ============================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                int m_blk, n_blk; tile_coords(tile, m_blk, n_blk);
                const int row_a = m_blk * GEMM2_BM + (int)cta_rank * GEMM_BM;
                const int row_b = n_blk * GEMM2_BN + (int)cta_rank * (GEMM2_BN / 2);
                for (int k = 0; k < p.num_k_iters; ++k) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::kStageBytes;
                    uint8_t* sb = sa + S::kABytes;
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);     // bytes of both CTAs
                    const uint32_t fb = mapa_cluster(smem_u32(&full_bar[stage]), 0);
                    tma_load_2d_pair(sa, &tmap_a, fb, k * GEMM_BK, row_a);
                    tma_load_2d_pair(sb, &tmap_b, fb, k * GEMM_BK, row_b);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer (leader CTA) ============================
        if (leader && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(GEMM2_BM, GEMM2_BN, false);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * GEMM2_BN;
                for (int k = 0; k < p.num_k_iters; ++k) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
                    const uint32_t sb = sa + S::kABytes;
                    #pragma unroll
                    for (int kk = 0; kk < GEMM_BK / 16; ++kk)
                        umma_bf16_ss_pair(d_tmem, umma_desc_kmajor_sw128(sa + kk * 32), umma_desc_kmajor_sw128(sb + kk * 32), idesc,
                                          (k | kk) != 0);
                    umma_commit_pair(&empty_bar[stage], 0b11);            // frees the slot in both CTAs when the MMAs retire
                    if (k == p.num_k_iters - 1) umma_commit_pair(&tfull_bar[acc], 0b11);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================ epilogue (both CTAs, own 128 rows) ============================
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
            int m_blk, n_blk; tile_coords(tile, m_blk, n_blk);
            const int m_cta = m_blk * GEMM2_BM + (int)cta_rank * GEMM_BM;      // first row of this CTA's half
            const long long m = (long long)m_cta + row;
            const bool row_ok = m < p.M;
            const long long row_off = m * p.ldc;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * GEMM2_BN + ((uint32_t)(wq * 32) << 16);
            if constexpr (EPI_TMA) {
                const bool lead_thr = (warp == 4 && lane == 0);
                #pragma unroll 1
                for (int c = 0; c < GEMM2_BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    const int n0 = n_blk * GEMM2_BN + c * 32;
                    float f[32];
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (n0 < p.N) {
                        if (p.bias) {
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                            }
                        }
                        if (p.gate) {
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + n0 + j));
                                f[j] *= g.x; f[j + 1] *= g.y; f[j + 2] *= g.z; f[j + 3] *= g.w;
                            }
                        }
                    }
                    uint8_t* buf = epi_smem + (c & 1) * (GEMM_BM * 128);
                    if (lead_thr) bulk_wait_group_read<1>();          // the reduce issued two chunks ago has read this buffer
                    named_bar_sync(1, 128);
                    #pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(buf + row * 128 + ((j ^ (row & 7)) << 4)) =
                            make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (lead_thr && n0 < p.N && m_cta < p.M) {
                        tma_reduce_add_2d(&tmap_c, buf, n0, m_cta);      // rows >= M are clipped by the tensor map
                        bulk_commit_group();
                    }
                }
            } else {
                #pragma unroll 1
                for (int c = 0; c < GEMM2_BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    const int n0 = n_blk * GEMM2_BN + c * 32;
                    if (row_ok && n0 < p.N) {
                        float f[32];
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        const int ncols = min(32, p.N - n0);          // N % 8 == 0
                        if (p.bias) {
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (j < ncols) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                                    f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                                }
                            }
                        }
                        if (p.act == ACT_GELU_TANH) {
                            #pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = gelu_tanh(f[j]);
                        } else if (p.act == ACT_SILU) {
                            #pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = f[j] / (1.f + __expf(-f[j]));
                        } else if (p.act == ACT_GELU_ERF) {
                            #pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.7071067811865476f));
                        }
                        if (p.gate) {
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (j < ncols) {
                                    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + n0 + j));
                                    f[j] *= g.x; f[j + 1] *= g.y; f[j + 2] *= g.z; f[j + 3] *= g.w;
                                }
                            }
                        }
                        const long long off = row_off + n0;
                        if (p.residual) {
                            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off);
                            #pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                if (j < ncols) {
                                    const uint4 r4 = __ldg(rp + j / 8);
                                    const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
                                    #pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        f[j + 2 * q] += __uint_as_float(rw[q] << 16);
                                        f[j + 2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
                                    }
                                }
                            }
                        }
                        if (p.out_fp32) {
                            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (j < ncols) {
                                    float4 o = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                                    if (p.accumulate) {
                                        const float4 old = op[j / 4];
                                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                    }
                                    op[j / 4] = o;
                                }
                            }
                        } else {
                            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
                            #pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                if (j < ncols) {
                                    uint4 o;
                                    o.x = pack_bf16x2(f[j], f[j + 1]); o.y = pack_bf16x2(f[j + 2], f[j + 3]);
                                    o.z = pack_bf16x2(f[j + 4], f[j + 5]); o.w = pack_bf16x2(f[j + 6], f[j + 7]);
                                    op[j / 8] = o;
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // this CTA's accumulator rows have been read: one arrival per epilogue warp on the LEADER's barrier
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_cluster(smem_u32(&tempty_bar[acc]), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    if (EPI_TMA && warp == 4 && lane == 0) bulk_wait_group<0>();   // smem staging tiles must outlive the reduce-stores
    tc_fence_before();
    cluster_sync_all();               // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers / TMEM
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace b200
