// Row-tiled implicit-GEMM convolution for sm_100a: the activation halo is staged ONCE per (frame tap, channel chunk)
// and all kh x kw spatial taps are issued from the same shared-memory tile through shifted UMMA descriptors.
//
// gemm_sm100.cuh's MODE_CONV loads one 16 KB A box per tap: at Cout <= 128 the 27 boxes per chunk made the kernel
// L2->SMEM bound (ncu launch list, profiles/: the 128->128 and the 128->3 full-resolution convs of the Hunyuan VAEs take the
// same 33 ms per 34-frame slice, i.e. the time is the A traffic of 216 GB per launch, not the MMAs).  Here the M tile is
// 128 CONSECUTIVE pixels of one image row (x ROWS rows), so the A tile of tap (dh, dw) is the same smem box shifted by
// (dh * WB + dw) rows of 128 B: a legal K-major SWIZZLE_128B operand whose start address is not 1024-B aligned.  The 128B
// swizzle of both the TMA write and the tcgen05.mma read is a function of the ABSOLUTE smem address bits, so the shifted
// descriptor needs no fix-up: measured on B200, parity holds with the base-offset field 0 and breaks when it is set to
// (addr >> 7) & 7.  A traffic drops 9x (3x3 taps) / (halo overhead 1.02-2x), and
// with ROWS = 2 each weight tile B feeds two accumulators.
//
//   warp 0   TMA producer: per (dt, chunk) one 4-D box {64 ch, 128+kw-1, ROWS+kh-1, 1}; per tap one weight box {64, 1, BN}
//   warp 1   MMA issuer  : per tap ROWS x 4 tcgen05.mma (M=128, N=BN, K=16) into ROWS accumulators
//   warp 2   TMEM allocator (2 x 256 columns: epilogue of tile i overlaps the main loop of tile i+1)
//   warps 4-7 epilogue (bias / residual / bf16, interleaved or planar-fp32 stores; same contract as gemm_sm100.cuh)
#pragma once
#include <cuda.h>

#include "gemm_sm100.cuh"

namespace b200 {

constexpr int CONVR_BW = 128;                 // M tile = 128 consecutive pixels of an image row

// BKC = channels per K chunk: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B: Cin = 96 = 3 x 32 exactly,
// no zero-padded K is multiplied)
// PAIR: two CTAs of a cluster work on two adjacent pixel tiles and SHARE every weight tile (tcgen05.mma.cta_group::2, M = 256): each
// stages and reads half of the Cout rows of B.
template <int BN, int ROWS, int BKC = 64, bool PAIR = false>
struct ConvRowSmem {
    static_assert(ROWS * BN <= 256, "ROWS accumulators of BN columns per TMEM buffer");
    static_assert(BKC == 64 || BKC == 32, "K chunk");
    static_assert(!PAIR || BN % 32 == 0, "pair: each CTA stages BN/2 weight rows, UMMA N % 16 == 0");
    static constexpr int kRowBytes = BKC * 2;
    static constexpr int kAStage = (((CONVR_BW + 2) * (ROWS + 2) * kRowBytes) + 1023) / 1024 * 1024;     // sized for 3x3 taps
    static constexpr int kBRows = PAIR ? BN / 2 : BN;
    static constexpr int kBStage = (kBRows * kRowBytes + 1023) / 1024 * 1024;
    static constexpr int kBBytes = kBRows * kRowBytes;
    static constexpr int kAStages = 2;
    static constexpr int kBMax = (200 * 1024 - kAStages * kAStage) / kBStage;
    static constexpr int kBStages = kBMax > 8 ? 8 : kBMax;
    static_assert(kBStages >= 2, "weight ring too small");
    static constexpr int kBytes = kAStages * kAStage + kBStages * kBStage + 1024 /*align*/ + 256 /*barriers*/;
};

// K-major SWIZZLE_128B operand whose first row is NOT at a 1024-B boundary: rows stay 128 B apart (SBO = 1024 per 8 rows).
// use_base_offset (bits 49-51 = swizzle phase of the first row) exists for the A/B experiment only; the correct value is 0.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128_rowoff(uint32_t smem_addr, int use_base_offset) {
    uint64_t d = umma_desc_kmajor_sw128(smem_addr);
    if (use_base_offset) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
    return d;
}

// NORM: the epilogue also applies the RMS_norm (+SiLU) of the layer that consumes this conv's output (models/wan/modules/vae.py:85-103,
// 246-250: F.normalize over channels * sqrt(C) * gamma, then SiLU) -- every epilogue thread owns one pixel and, with a single N tile,
// all of its channels, so the statistics need no exchange.  The separate norm pass (one read + one write of every activation,
// 15 % of the Wan decode) disappears: a ResidualBlock's first conv writes ONLY the normalised tensor, its second conv (and the
// up-sampling convs) write the raw tensor for the skip path plus the normalised one for the next block.
// PAIR (launched with a cluster dimension of 2): CTA r of a pair owns pixel tile 2 * pair + r and weight rows [r BN/2, (r+1) BN/2); the
// leader issues M = 256 MMAs that read the activation halo tile of each CTA at the same shared-memory offset and both halves of the
// weight tile.  Why: at Cout = 96 / 192 the single-CTA kernel is bound by the shared-memory pipe, not the tensor pipe -- per
// M128 x N192 x K16 MMA (96 clk) it reads 4 KB of A + 6 KB of B and the TMA refills ~7.5 KB (B is re-fetched for every tile), 187 B/clk
// against 128 B/clk; with the weight tile shared the same MMA costs 4 + 3 KB of reads and ~4.5 KB of refills.
template <int BN, int ROWS, int BKC = 64, bool NORM = false, bool PAIR = false>
__global__ void __launch_bounds__(256, 1)
conv_row_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using S = ConvRowSmem<BN, ROWS, BKC, PAIR>;
    constexpr int RB = S::kRowBytes;                          // bytes per pixel row of the smem tiles
    constexpr int SA = S::kAStages, SB = S::kBStages;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");

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

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        // PAIR: "accumulator empty" lives on the leader: one arrival per epilogue warp of each CTA
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], PAIR ? 8 : 128); }
        fence_mbar_init();
    }
    if (warp == 2) { if constexpr (PAIR) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // PAIR: the persistent loop runs over PAIRS of pixel tiles (n_tiles == 1); this CTA's tile is 2 * pair + rank
    const int num_tiles = PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles * p.n_tiles;
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int taps_hw = p.kh * p.kw;
    const int WB = CONVR_BW + p.kw - 1;                       // box width (pixels) = smem rows per image row
    // tile -> (n_blk fastest, then w tile, then frame, then h band): the frames a causal conv re-reads stay in L2
    auto tile_coords = [&](int tile, int& n_blk, int& t0, int& h0, int& w0) {
        int m_blk = tile / p.n_tiles;
        n_blk = tile - m_blk * p.n_tiles;
        if constexpr (PAIR) { m_blk = 2 * tile + (int)cta_rank; n_blk = 0; }      // an odd tile count leaves one all-out-of-range tile (h0 >= H)
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
            const uint32_t a_bytes = (uint32_t)WB * (ROWS + p.kh - 1) * RB;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int n_blk, t0, h0, w0; tile_coords(tile, n_blk, t0, h0, w0);
                for (int dt = 0; dt < p.kt; ++dt) {
                    for (int cc = 0; cc < p.cin_chunks; ++cc) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        // causal in time (all padding in front), centred in space; OOB -> zero fill
                        if constexpr (PAIR) {       // both CTAs' bytes are credited to the leader's barrier
                            if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * a_bytes);
                            tma_load_4d_pair(smem_a + sa * S::kAStage, &tmap_a, mapa_cluster(smem_u32(&a_full[sa]), 0), cc * BKC, w0 - p.pad_w,
                                             h0 - p.pad_h, t0 + dt - p.pad_t);
                        } else {
                            mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                            tma_load_4d(smem_a + sa * S::kAStage, &tmap_a, &a_full[sa], cc * BKC, w0 - p.pad_w, h0 - p.pad_h, t0 + dt - p.pad_t);
                        }
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                        for (int tap = 0; tap < taps_hw; ++tap) {
                            mbar_wait(&b_empty[sb], pb ^ 1);
                            if constexpr (PAIR) {   // this CTA's half of the Cout rows
                                if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * S::kBBytes);
                                tma_load_3d_pair(smem_b + sb * S::kBStage, &tmap_b, mapa_cluster(smem_u32(&b_full[sb]), 0), cc * BKC, dt * taps_hw + tap,
                                                 (int)cta_rank * (BN / 2));
                            } else {
                                mbar_arrive_expect_tx(&b_full[sb], S::kBBytes);
                                tma_load_3d(smem_b + sb * S::kBStage, &tmap_b, &b_full[sb], cc * BKC, dt * taps_hw + tap, n_blk * BN);
                            }
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
            constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * GEMM_BM : GEMM_BM, BN, false);
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const int use_bo = p.conv_base_offset;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                bool first = true;
                for (int dt = 0; dt < p.kt; ++dt) {
                    for (int cc = 0; cc < p.cin_chunks; ++cc) {
                        mbar_wait(&a_full[sa], pa);
                        tc_fence_after();
                        const uint32_t a_base = smem_u32(smem_a + sa * S::kAStage);
                        for (int dh = 0; dh < p.kh; ++dh) {
                            for (int dw = 0; dw < p.kw; ++dw) {
                                mbar_wait(&b_full[sb], pb);
                                tc_fence_after();
                                const uint32_t b_base = smem_u32(smem_b + sb * S::kBStage);
                                #pragma unroll
                                for (int r = 0; r < ROWS; ++r) {
                                    const uint32_t a_row = a_base + (uint32_t)((r + dh) * WB + dw) * RB;
                                    #pragma unroll
                                    for (int kk = 0; kk < BKC / 16; ++kk) {
                                        const uint64_t da = BKC == 64 ? umma_desc_kmajor_sw128_rowoff(a_row + kk * 32, use_bo)
                                                                      : umma_desc_kmajor_sw64(a_row + kk * 32);
                                        const uint64_t db = BKC == 64 ? umma_desc_kmajor_sw128(b_base + kk * 32) : umma_desc_kmajor_sw64(b_base + kk * 32);
                                        if constexpr (PAIR) umma_bf16_ss_pair(d_tmem + r * BN, da, db, idesc, !(first && kk == 0));
                                        else umma_bf16_ss(d_tmem + r * BN, da, db, idesc, !(first && kk == 0));
                                    }
                                }
                                first = false;
                                if constexpr (PAIR) umma_commit_pair(&b_empty[sb], 0b11); else umma_commit(&b_empty[sb]);   // frees the weight slot
                                if (++sb == SB) { sb = 0; pb ^= 1; }
                            }
                        }
                        if constexpr (PAIR) umma_commit_pair(&a_empty[sa], 0b11); else umma_commit(&a_empty[sa]);           // frees the halo tile
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
                }
                if constexpr (PAIR) umma_commit_pair(&tfull_bar[acc], 0b11); else umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================ epilogue ============================
        const int wq = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = wq * 32 + lane;             // accumulator row = pixel of the row tile
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            int n_blk, t0, h0, w0; tile_coords(tile, n_blk, t0, h0, w0);
            const int w = w0 + row;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            #pragma unroll 1
            for (int r = 0; r < ROWS; ++r) {
                const int h = h0 + r;
                const bool row_ok = (h < p.H) && (w < p.W);
                const long long row_off = t0 * p.st_t + h * p.st_h + w * p.st_w;
                const uint32_t t_row = tmem_base + acc * 256 + r * BN + ((uint32_t)(wq * 32) << 16);
                if constexpr (NORM) {
                    // bf16 channels-last output, one N tile (BN == N, BN % 32 == 0): keep the rounded row packed in registers
                    static_assert(BN % 32 == 0, "NORM epilogue: whole 32-column chunks");
                    uint32_t pk[BN / 2];
                    float ss = 0.f;
                    #pragma unroll
                    for (int c = 0; c < BN / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(t_row + c * 32, v);
                        tmem_ld_wait();
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
                        if (p.residual && row_ok) {
                            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + c * 32);
                            #pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                const uint4 r4 = __ldg(rp + j / 8);
                                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
                                #pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    f[j + 2 * q] += __uint_as_float(rw[q] << 16);
                                    f[j + 2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
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
                    const float inv = sqrtf((float)BN) / fmaxf(sqrtf(ss), 1e-12f);
                    if (row_ok) {
                        uint4* np = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.norm_out) + row_off);
                        #pragma unroll
                        for (int q = 0; q < BN / 8; ++q) {
                            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.norm_gamma + q * 8));
                            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.norm_gamma + q * 8 + 4));
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
                    continue;
                }
                #pragma unroll 1
                for (int c = 0; c < BN / 32 + (BN % 32 ? 1 : 0); ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    const int n0 = n_blk * BN + c * 32;
                    if (row_ok && n0 < p.N) {
                        float f[32];
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        const int ncols = min(min(32, BN - c * 32), p.N - n0);
                        if (p.planar) {
                            // few-channel planar fp32 output (VAE head, Cout = 3): out[n][pixel]
                            #pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols)
                                    reinterpret_cast<float*>(p.out)[(long long)(n0 + j) * p.st_split + row_off] = f[j] + (p.bias ? __ldg(p.bias + n0 + j) : 0.f);
                        } else {
                            if (p.bias) {
                                #pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    if (j < ncols) {
                                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                                        f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                                    }
                                }
                            }
                            long long off = row_off;
                            if (p.csplit > 0) off += (long long)(n0 / p.csplit) * p.st_split + (n0 % p.csplit);
                            else off += n0;
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
                            if (p.out_fp32) {                   // fp32 channels-last (latent moments of the Hunyuan VAE encoder)
                                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
                                #pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    if (j < ncols) op[j / 4] = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
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
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            if constexpr (PAIR) {
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(mapa_cluster(smem_u32(&tempty_bar[acc]), 0));
            } else {
                mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
