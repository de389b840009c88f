// Row-tiled implicit-GEMM convolution, second generation (CTA pairs only): conv_sm100.cuh's kernel with the three things its ncu
// capture (profiles/ncu_r02_conv.txt + the per-instruction samples of the same report) showed to be in the way at Cout = 96 / 192:
//
//  1. ISSUE-BOUND MMA WARP.  conv_row<96,2,32> issues 4 tcgen05.mma (192 tensor clocks) per tap, and one tap iteration of the issuing
//     thread -- mbarrier try-wait, runtime (dh, dw) address arithmetic, two descriptor rebuilds per MMA, commit -- is 58 instructions
//     = ~180 clocks of a single thread: the tensor pipe (55 % busy) waited for its instruction stream, not for operands (shared-memory
//     operand wavefronts 50 %, L2 47 %, DRAM 16 %).  Here kh x kw is a template parameter: the taps are fully unrolled, every operand
//     descriptor is "base descriptor + compile-time constant", and the weights of GROUP taps travel as ONE pipeline stage, so the
//     issuing thread waits once per GROUP taps (GROUP = kh*kw: one wait per 36 MMAs at Cin-chunk 32).
//  2. EPILOGUE AT 73 % OF THE TILE TIME with one warp per scheduler (no latency hiding; ~9 clk per instruction).  Two epilogue
//     warpgroups: with ROWS = 2 each owns one image row of the tile, with ROWS = 1 each owns half of the channels (the RMS statistics
//     are exchanged through shared memory).
//  3. The "accumulator drained" arrival on the leader's barrier was mbarrier.arrive.release.cluster = MEMBAR.ALL.GPU + ERRBAR per
//     arrival (12 % of the epilogue warps' samples).  The arrival only has to follow this thread's TMEM reads (tcgen05.wait::ld +
//     tcgen05.fence::before_thread_sync order those); no global / shared data is published, so it is a relaxed arrive.
//
//   warp 0      TMA producer: per (dt, chunk) one 4-D halo box; per GROUP of taps GROUP weight boxes on one barrier
//   warp 1      MMA issuer (leader CTA): unrolled taps, M = 256 (cta_group::2), weights shared by the two pixel tiles of the pair
//   warp 2      TMEM allocator (2 x 256 columns)
//   warps 4-11  two epilogue warpgroups (bias / residual / bf16 [/ fused next-layer RMS_norm + SiLU], same contract as conv_sm100.cuh)
#pragma once
#include <cuda.h>

#include "conv_sm100.cuh"

namespace b200 {

constexpr int CONV2_THREADS = 384;

template <int BN, int ROWS, int BKC, int KHW, int GROUP>
struct ConvRow2Smem {
    static_assert(ROWS * BN <= 256, "ROWS accumulators of BN columns per TMEM buffer");
    static_assert(BKC == 64 || BKC == 32, "K chunk");
    static_assert(KHW == 2 || KHW == 3, "spatial taps: 3x3, or the 2x2 parity convs of the folded 2x up-sampling");
    static_assert(BN % 32 == 0, "each CTA stages BN/2 weight rows, UMMA N % 16 == 0");
    static constexpr int kTaps = KHW * KHW;
    static_assert(kTaps % GROUP == 0, "taps per weight stage");
    static constexpr int kRowBytes = BKC * 2;
    static constexpr int kWB = CONVR_BW + KHW - 1;                 // halo box width (pixels) = smem rows per image row
    static constexpr int kABox = kWB * (ROWS + KHW - 1) * kRowBytes;
    static constexpr int kAStage = (kABox + 1023) / 1024 * 1024;
    static constexpr int kBRows = BN / 2;
    static constexpr int kBBytes = kBRows * kRowBytes;
    static constexpr int kBTap = (kBBytes + 1023) / 1024 * 1024;
    static constexpr int kBStage = GROUP * kBTap;
    static constexpr int kAStages = BKC == 32 ? 3 : 2;
    static constexpr int kBMax = (214 * 1024 - kAStages * kAStage) / kBStage;
    static constexpr int kBStages = kBMax > 4 ? 4 : kBMax;
    static_assert(kBStages >= 2, "weight ring too small");
    static constexpr int kBytes = kAStages * kAStage + kBStages * kBStage + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*RMS exchange*/;
    static_assert(kBytes <= 227 * 1024, "shared memory");
};

template <int BN, int ROWS, int BKC, bool NORM, int KHW, int GROUP>
__global__ void __launch_bounds__(CONV2_THREADS, 1)
conv_row2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using S = ConvRow2Smem<BN, ROWS, BKC, KHW, GROUP>;
    constexpr int RB = S::kRowBytes;
    constexpr int SA = S::kAStages, SB = S::kBStages;
    constexpr int WB = S::kWB;
    constexpr int NG = S::kTaps / GROUP;
    static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "UMMA N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + SA * S::kAStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + SB * S::kBStage);
    uint64_t* a_full = bars;                 // [SA]
    uint64_t* a_empty = bars + SA;           // [SA]
    uint64_t* b_full = bars + 2 * SA;        // [SB]
    uint64_t* b_empty = bars + 2 * SA + SB;  // [SB]
    uint64_t* tfull_bar = bars + 2 * SA + 2 * SB;       // [2]
    uint64_t* tempty_bar = bars + 2 * SA + 2 * SB + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 4);
    float* ss_x = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);      // [2 tiles][2 warpgroups][128 pixels]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        // "accumulator empty" lives on the leader: one arrival per epilogue warp (8) of each CTA
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 16); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // persistent loop over PAIRS of pixel tiles; this CTA's tile is 2 * pair + rank
    const int num_tiles = (p.m_tiles + 1) / 2;
    const int tile0 = (int)(blockIdx.x >> 1);
    const int tile_step = (int)(gridDim.x >> 1);
    // pixel tile -> (w tile fastest, then frame, then h band): the frames a causal conv re-reads stay in L2
    auto tile_coords = [&](int tile, int& t0, int& h0, int& w0) {
        const int m_blk = 2 * tile + (int)cta_rank;          // an odd tile count leaves one all-out-of-range tile (h0 >= H)
        const int per_band = p.T * p.tiles_w;
        const int band = m_blk / per_band;
        const int r = m_blk - band * per_band;
        t0 = r / p.tiles_w;
        h0 = band * ROWS;
        w0 = (r - t0 * p.tiles_w) * CONVR_BW;
    };

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            const uint32_t afull0 = mapa_cluster(smem_u32(&a_full[0]), 0), bfull0 = mapa_cluster(smem_u32(&b_full[0]), 0);
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int t0, h0, w0; tile_coords(tile, t0, h0, w0);
                for (int dt = 0; dt < p.kt; ++dt) {
                    for (int cc = 0; cc < p.cin_chunks; ++cc) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        // causal in time (all padding in front), centred in space; OOB -> zero fill.  Both CTAs' bytes are credited to
                        // the leader's barrier.
                        if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * S::kABox);
                        tma_load_4d_pair(smem_a + sa * S::kAStage, &tmap_a, afull0 + sa * 8, cc * BKC, w0 - p.pad_w, h0 - p.pad_h, t0 + dt - p.pad_t);
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                        #pragma unroll 1
                        for (int g = 0; g < NG; ++g) {
                            mbar_wait(&b_empty[sb], pb ^ 1);
                            if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * GROUP * S::kBBytes);
                            #pragma unroll
                            for (int t = 0; t < GROUP; ++t)      // this CTA's half of the Cout rows of tap g * GROUP + t
                                tma_load_3d_pair(smem_b + sb * S::kBStage + t * S::kBTap, &tmap_b, bfull0 + sb * 8, cc * BKC,
                                                 dt * S::kTaps + g * GROUP + t, (int)cta_rank * (BN / 2));
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (leader && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN, false);
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                uint32_t accum0 = 0;                              // the very first MMA of each accumulator overwrites
                for (int dt = 0; dt < p.kt; ++dt) {
                    for (int cc = 0; cc < p.cin_chunks; ++cc) {
                        mbar_wait(&a_full[sa], pa);
                        tc_fence_after();
                        const uint32_t a_base = smem_u32(smem_a + sa * S::kAStage);
                        const uint64_t da0 = BKC == 64 ? umma_desc_kmajor_sw128(a_base) : umma_desc_kmajor_sw64(a_base);
                        #pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            mbar_wait(&b_full[sb], pb);
                            tc_fence_after();
                            const uint32_t b_base = smem_u32(smem_b + sb * S::kBStage);
                            const uint64_t db0 = BKC == 64 ? umma_desc_kmajor_sw128(b_base) : umma_desc_kmajor_sw64(b_base);
                            #pragma unroll
                            for (int t = 0; t < GROUP; ++t) {
                                const int tap = g * GROUP + t, dh = tap / KHW, dw = tap % KHW;     // compile-time after unrolling
                                #pragma unroll
                                for (int r = 0; r < ROWS; ++r) {
                                    #pragma unroll
                                    for (int kk = 0; kk < BKC / 16; ++kk) {
                                        // the start-address field holds (addr >> 4) & 0x3fff and shared memory ends below 2^18: adding a byte
                                        // offset >> 4 cannot carry out of the field
                                        const uint64_t da = da0 + (uint64_t)((((r + dh) * WB + dw) * RB + kk * 32) >> 4);
                                        const uint64_t db = db0 + (uint64_t)((t * S::kBTap + kk * 32) >> 4);
                                        umma_bf16_ss_pair(d_tmem + r * BN, da, db, idesc, (tap == 0 && kk == 0) ? accum0 : 1u);
                                    }
                                }
                            }
                            umma_commit_pair(&b_empty[sb], 0b11);       // frees the weight stage in both CTAs
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                        accum0 = 1;
                        umma_commit_pair(&a_empty[sa], 0b11);           // frees the halo tile in both CTAs
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
                }
                umma_commit_pair(&tfull_bar[acc], 0b11);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================ epilogue (two warpgroups) ============================
        const int wg = (warp - 4) >> 2;
        const int wq = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = wq * 32 + lane;             // accumulator row = pixel of the row tile
        const uint32_t tempty0 = mapa_cluster(smem_u32(&tempty_bar[0]), 0);
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            int t0, h0, w0; tile_coords(tile, t0, h0, w0);
            const int w = w0 + row;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if constexpr (NORM) {
                // bf16 channels-last output, one N tile (BN == N).  ROWS == 2: warpgroup wg owns image row wg of the tile and all BN channels;
                // ROWS == 1: it owns channels [wg BN/2, (wg+1) BN/2) and the two partial sums of squares meet in shared memory.
                constexpr int NC = ROWS == 2 ? BN : BN / 2;
                static_assert(ROWS <= 2 && NC % 32 == 0, "NORM epilogue: whole 32-column chunks per warpgroup");
                const int r = ROWS == 2 ? wg : 0;
                const int col0 = ROWS == 2 ? 0 : wg * NC;
                const int h = h0 + r;
                const bool row_ok = (h < p.H) && (w < p.W);
                const long long row_off = t0 * p.st_t + h * p.st_h + w * p.st_w + col0;
                const uint32_t t_row = tmem_base + acc * 256 + r * BN + col0 + ((uint32_t)(wq * 32) << 16);
                uint32_t pk[NC / 2];
                float ss = 0.f;
                #pragma unroll
                for (int c = 0; c < NC / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    uint4 r4[4];                                  // the skip tensor's 32 channels: in flight while TMEM is read
                    if (p.residual && row_ok) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + c * 32);
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) r4[j] = __ldg(rp + j);
                    }
                    tmem_ld_wait();
                    float f[32];
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.bias) {
                        #pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c * 32 + j));
                            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                        }
                    }
                    if (p.residual && row_ok) {
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t rw[4] = {r4[j].x, r4[j].y, r4[j].z, r4[j].w};
                            #pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                f[8 * j + 2 * q] += __uint_as_float(rw[q] << 16);
                                f[8 * j + 2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
                            }
                        }
                    }
                    #pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t w2 = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                        pk[c * 16 + j] = w2;
                        const float a = __uint_as_float(w2 << 16), b = __uint_as_float(w2 & 0xffff0000u);   // the ROUNDED values, as the
                        ss = fmaf(a, a, fmaf(b, b, ss));                                                    // separate norm pass saw them
                    }
                    if (!p.norm_only && row_ok) {
                        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + c * 32);
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = make_uint4(pk[c * 16 + 4 * j], pk[c * 16 + 4 * j + 1], pk[c * 16 + 4 * j + 2], pk[c * 16 + 4 * j + 3]);
                    }
                }
                // the accumulator is in registers: hand the TMEM buffer back before the normalisation arithmetic
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + acc * 8);
                if constexpr (ROWS == 1) {
                    float* sx = ss_x + acc * 256;
                    sx[wg * 128 + row] = ss;
                    named_bar_sync(1, 256);                       // the 8 epilogue warps
                    ss += sx[(wg ^ 1) * 128 + row];
                }
                const float inv = sqrtf((float)BN) / fmaxf(sqrtf(ss), 1e-12f);
                if (row_ok) {
                    uint4* np = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.norm_out) + row_off);
                    #pragma unroll
                    for (int q = 0; q < NC / 8; ++q) {
                        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.norm_gamma + col0 + q * 8));
                        const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.norm_gamma + col0 + q * 8 + 4));
                        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                        uint32_t o[4];
                        #pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t w2 = pk[q * 4 + k];
                            const float a = silu_fast(__uint_as_float(w2 << 16) * inv * g[2 * k]);
                            const float b = silu_fast(__uint_as_float(w2 & 0xffff0000u) * inv * g[2 * k + 1]);
                            o[k] = pack_bf16x2(a, b);
                        }
                        np[q] = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                __syncwarp();
            } else {
                // plain epilogue: bias / residual -> bf16 or fp32 channels-last; the (row, 32-column chunk) units alternate between the warpgroups
                constexpr int NCH = BN / 32;
                static_assert(BN % 32 == 0, "whole 32-column chunks");
                #pragma unroll 1
                for (int u = wg; u < ROWS * NCH; u += 2) {
                    const int r = u / NCH, c = u - r * NCH;
                    const int h = h0 + r;
                    const bool row_ok = (h < p.H) && (w < p.W);
                    const long long off = t0 * p.st_t + h * p.st_h + w * p.st_w + c * 32;
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + acc * 256 + r * BN + c * 32 + ((uint32_t)(wq * 32) << 16), v);
                    uint4 r4[4];
                    if (p.residual && row_ok) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off);
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) r4[j] = __ldg(rp + j);
                    }
                    tmem_ld_wait();
                    if (row_ok) {
                        float f[32];
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (p.bias) {
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c * 32 + j));
                                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                            }
                        }
                        if (p.residual) {
                            #pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t rw[4] = {r4[j].x, r4[j].y, r4[j].z, r4[j].w};
                                #pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    f[8 * j + 2 * q] += __uint_as_float(rw[q] << 16);
                                    f[8 * j + 2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
                                }
                            }
                        }
                        if (p.out_fp32) {                       // fp32 channels-last (latent moments of the Hunyuan VAE encoder)
                            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) op[j / 4] = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        } else {
                            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
                            #pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 o;
                                o.x = pack_bf16x2(f[j], f[j + 1]); o.y = pack_bf16x2(f[j + 2], f[j + 3]);
                                o.z = pack_bf16x2(f[j + 4], f[j + 5]); o.w = pack_bf16x2(f[j + 6], f[j + 7]);
                                op[j / 8] = o;
                            }
                        }
                    }
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + acc * 8);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace b200
