// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T + bias )         bf16 operands, fp32 accumulate in TMEM
//
// One CTA per SM (grid = min(#tiles, #SMs)), 256 threads:
//   warp 0   TMA producer   (one elected lane; cp.async.bulk.tensor into a BK=64, 128B-swizzled smem ring)
//   warp 1   MMA issuer     (one elected lane; tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction)
//   warp 2   TMEM allocator (512 columns = 2 accumulator buffers, so the epilogue of tile i overlaps
//                            the main loop of tile i+1)
//   warps 4-7 epilogue      (tcgen05.ld 32x32b: thread <-> accumulator row; bias / GELU / gate / residual
//                            fused; 16-byte global stores)
//
// MODE_CONV turns the same main loop into an implicit-GEMM causal convolution over a channels-last
// activation [T,H,W,C]: the M tile is an 8x16 pixel patch of one frame, the K loop runs over
// (tap, 64-channel chunk) and the A tile of each step is ONE 4-D TMA box at the tap-shifted
// coordinate.  Spatial zero padding and the causal temporal zero padding are the TMA's out-of-bounds
// zero fill, so no padded copy of the activation ever exists (reference: models/wan/modules/vae.py:43-63
// materialises F.pad / torch.cat copies before cuDNN).
#pragma once
#include <cuda.h>

#include "sm100.cuh"

namespace b200 {

enum { MODE_LINEAR = 0, MODE_CONV = 1 };
enum { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_SILU = 2, ACT_GELU_ERF = 3 };

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int CONV_BH = 8;    // M tile = CONV_BH x CONV_BW pixels of one frame
constexpr int CONV_BW = 16;

struct GemmParams {
    int M, N, K;              // linear: problem size.  conv: N = Cout, K unused
    int mode;
    int num_k_iters;          // linear: ceil(K/64).  conv: taps * cin_chunks
    // ---- conv geometry (input and output share the T,H,W pixel grid)
    int T, H, W;
    int kt, kh, kw;           // taps
    int pad_h, pad_w;         // taps start at (h - pad_h, w - pad_w): kh/2, kw/2 for centred convs
    int pad_t;                // taps start at t - pad_t: kt-1 (causal); 0 when the input was padded explicitly (replicate)
    int cin_chunks;           // ceil(Cin / 64)
    int tiles_h, tiles_w;     // ceil(H/8), ceil(W/16)
    // ---- tile rasterisation
    int m_tiles, n_tiles, n_group;   // n tiles are walked in groups of n_group (B stays L2 resident)
    // ---- epilogue
    void* out;                // bf16 or fp32
    int out_fp32;
    int accumulate;           // out (fp32) += value   (DiT residual stream)
    int tma_reduce;           // accumulate through cp.reduce.async.bulk.tensor (EPI_TMA instantiation)
    long long ldc;            // linear: row stride (elements)
    long long st_t, st_h, st_w;   // conv: pixel strides (elements)
    int csplit;               // conv: columns >= csplit go to a second plane (time_conv interleave)
    long long st_split;
    int planar;               // conv: out is fp32 [N][T*H*W] planes (st_split = plane stride, st_* = pixel strides)
    const float* bias;        // [N] or null
    const float* gate;        // [N] or null : value *= gate[n]
    const __nv_bfloat16* residual;   // same mapping as out (bf16) or null : value += residual
    int act;
    int conv_base_offset;     // conv_sm100.cuh: put (addr >> 7) & 7 into the A descriptors' base-offset field
    // ---- fused "next layer's norm" epilogue (conv_sm100.cuh, NORM instantiations; requires one N tile = all channels of a pixel):
    // norm_out[pixel, :] = silu( bf16(value) / max(||bf16(value)||_2, 1e-12) * sqrt(N) * norm_gamma ), same pixel strides as out;
    // norm_only: the raw tensor is not written (its only consumer was the norm)
    const float* norm_gamma;
    void* norm_out;
    int norm_only;
};

// BKC = K elements per TMA box (64 -> 128B swizzle, 32 -> 64B swizzle); NBOX boxes of A and of B form one pipeline stage
// EPI_TMA: 2 x 16 KB staging tiles for the TMA reduce-add epilogue (fp32 residual-stream accumulate)
template <int BN, int BKC = 64, int NBOX = 1, bool EPI_TMA = false>
struct GemmSmem {
    static constexpr int kABox = GEMM_BM * BKC * 2;
    static constexpr int kBBox = BN * BKC * 2;
    static constexpr int kABytes = NBOX * kABox;                   // 16 KB for the default 128 x 64 tile
    static constexpr int kBBytes = NBOX * kBBox;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kEpiBytes = EPI_TMA ? 2 * GEMM_BM * 128 : 0;
    static constexpr int kBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// EPI_TMA (linear mode, fp32 accumulate): the epilogue stages each 128 x 32 fp32 chunk in 128B-swizzled smem and issues
// cp.reduce.async.bulk.tensor (.add): the read-modify-write of the residual stream happens in L2, fully coalesced, instead
// of per-thread row-strided LDG/STG (which cost 30% of the GEMM: 4.37 ms vs 2.98 ms at M=75600, N=K=5120).
template <int BN, bool B_MN_MAJOR, int BKC = 64, int NBOX = 1, bool EPI_TMA = false>
__global__ void __launch_bounds__(256, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
    using S = GemmSmem<BN, BKC, NBOX, EPI_TMA>;
    static_assert(BKC == 64 || BKC == 32, "K box: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B) bf16 elements");
    static_assert(!(B_MN_MAJOR && (BKC != 64 || NBOX != 1)), "MN-major B only with the default K box");
    constexpr int BKS = BKC * NBOX;                  // K elements per pipeline stage
    constexpr int kStages = S::kStages;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
    static_assert(!B_MN_MAJOR || BN % 64 == 0, "MN-major B needs 64-wide slabs");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + kStages * S::kStageBytes;          // [2][128 rows][128 B], EPI_TMA only
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + S::kEpiBytes);
    uint64_t* full_bar = bars;                    // [kStages]
    uint64_t* empty_bar = bars + kStages;         // [kStages]
    uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
    uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.m_tiles * p.n_tiles;
    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        // groups of n_group consecutive n-tiles; inside a group n is fastest
        const int per_group = p.m_tiles * p.n_group;
        const int g = tile / per_group;
        const int r = tile - g * per_group;
        const int gw = min(p.n_group, p.n_tiles - g * p.n_group);
        m_blk = r / gw;
        n_blk = g * p.n_group + (r - m_blk * gw);
    };

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk; tile_coords(tile, m_blk, n_blk);
                int t0 = 0, h0 = 0, w0 = 0;
                if (p.mode == MODE_CONV) {
                    // tile order (h-band, t, w): the 3 frames a causal 3x3x3 tile reads were touched within the last few
                    // hundred tiles, so temporal taps hit L2 (a whole 720p frame at 96 ch is 177 MB > L2); frame-major order
                    // re-read the input 3x from HBM (ncu: 6.8 GB read for a 2.3 GB input)
                    const int per_band = p.T * p.tiles_w;
                    const int band = m_blk / per_band;
                    const int r = m_blk - band * per_band;
                    t0 = r / p.tiles_w;
                    h0 = band * CONV_BH;
                    w0 = (r % p.tiles_w) * CONV_BW;
                }
                for (int k = 0; k < p.num_k_iters; ++k) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::kStageBytes;
                    uint8_t* sb = sa + S::kABytes;
                    mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
                    if (p.mode == MODE_LINEAR) {
                        if constexpr (!B_MN_MAJOR) {
                            #pragma unroll
                            for (int b = 0; b < NBOX; ++b) {
                                tma_load_2d(sa + b * S::kABox, &tmap_a, &full_bar[stage], k * BKS + b * BKC, m_blk * GEMM_BM);
                                tma_load_2d(sb + b * S::kBBox, &tmap_b, &full_bar[stage], k * BKS + b * BKC, n_blk * BN);
                            }
                        } else {
                            tma_load_2d(sa, &tmap_a, &full_bar[stage], k * GEMM_BK, m_blk * GEMM_BM);
                            // B is [K][N] (N contiguous): one [64 k-rows][64 n] slab per 64 columns
                            #pragma unroll
                            for (int s = 0; s < BN / 64; ++s)
                                tma_load_2d(sb + s * (GEMM_BK * 128), &tmap_b, &full_bar[stage], n_blk * BN + s * 64, k * GEMM_BK);
                        }
                    } else {
                        const int tap = k / p.cin_chunks;
                        const int cc = k - tap * p.cin_chunks;
                        const int dt = tap / (p.kh * p.kw);
                        const int rr = tap - dt * (p.kh * p.kw);
                        const int dh = rr / p.kw;
                        const int dw = rr - dh * p.kw;
                        // causal in time (all padding in front), centred in space; OOB -> zero fill
                        #pragma unroll
                        for (int b = 0; b < NBOX; ++b) {
                            tma_load_4d(sa + b * S::kABox, &tmap_a, &full_bar[stage], cc * BKS + b * BKC, w0 + dw - p.pad_w,
                                        h0 + dh - p.pad_h, t0 + dt - p.pad_t);
                            tma_load_3d(sb + b * S::kBBox, &tmap_b, &full_bar[stage], cc * BKS + b * BKC, tap, n_blk * BN);
                        }
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, B_MN_MAJOR);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int k = 0; k < p.num_k_iters; ++k) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
                    const uint32_t sb = sa + S::kABytes;
                    #pragma unroll
                    for (int b = 0; b < NBOX; ++b) {
                        #pragma unroll
                        for (int kk = 0; kk < BKC / 16; ++kk) {
                            const uint32_t aa = sa + b * S::kABox + kk * 32, bb = sb + b * S::kBBox + kk * 32;
                            const uint64_t da = BKC == 64 ? umma_desc_kmajor_sw128(aa) : umma_desc_kmajor_sw64(aa);
                            const uint64_t db = B_MN_MAJOR ? umma_desc_mnmajor_sw128(sb + kk * 2048, GEMM_BK * 128)
                                                           : (BKC == 64 ? umma_desc_kmajor_sw128(bb) : umma_desc_kmajor_sw64(bb));
                            umma_bf16_ss(d_tmem, da, db, idesc, (k | b | kk) != 0);
                        }
                    }
                    umma_commit(&empty_bar[stage]);                 // frees the smem slot when the MMAs retire
                    if (k == p.num_k_iters - 1) umma_commit(&tfull_bar[acc]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================ epilogue ============================
        const int wq = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = wq * 32 + lane;             // accumulator row
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk; tile_coords(tile, m_blk, n_blk);
            bool row_ok;
            long long row_off;
            if (p.mode == MODE_LINEAR) {
                const long long m = (long long)m_blk * GEMM_BM + row;
                row_ok = m < p.M;
                row_off = m * p.ldc;
            } else {
                const int per_band = p.T * p.tiles_w;
                const int band = m_blk / per_band;
                const int r = m_blk - band * per_band;
                const int t0 = r / p.tiles_w;
                const int h = band * CONV_BH + row / CONV_BW;
                const int w = (r % p.tiles_w) * CONV_BW + row % CONV_BW;
                row_ok = (h < p.H) && (w < p.W);
                row_off = t0 * p.st_t + h * p.st_h + w * p.st_w;
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * 256 + ((uint32_t)(wq * 32) << 16);
            if constexpr (EPI_TMA) {
                const bool leader = (warp == 4 && lane == 0);
                #pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int n0 = n_blk * BN + c * 32;
                    // N tail (N % BN != 0): the remaining chunks lie beyond N.  They must not touch the staging buffers: a chunk that
                    // stages without committing a store breaks the "buffer (c & 1) was committed two chunks ago" accounting of
                    // bulk_wait_group_read<1> below and overwrites a tile the TMA engine may still be reading (found with the byT5
                    // widths, N = 1472: the last 32 valid columns raced).  n0 is uniform over the 128 epilogue threads.
                    if (n0 >= p.N) break;
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    float f[32];
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (n0 < p.N) {        // N % 32 == 0 is required by the host for this path
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
                    // buffer (c & 1) was handed to the TMA two chunks ago: wait until that store has READ it
                    if (leader) bulk_wait_group_read<1>();
                    named_bar_sync(1, 128);
                    #pragma unroll
                    for (int j = 0; j < 8; ++j)            // 128-B row, 16-B chunks XOR-swizzled by (row & 7): conflict-free
                        *reinterpret_cast<float4*>(buf + row * 128 + ((j ^ (row & 7)) << 4)) =
                            make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (leader && n0 < p.N) {
                        tma_reduce_add_2d(&tmap_c, buf, n0, m_blk * GEMM_BM);     // rows >= M are clipped by the tensor map
                        bulk_commit_group();
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
                // columns handled by this chunk (N is a multiple of 16, chunks of 32 may be half full)
                const int ncols = min(32, p.N - n0);
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
                long long off = row_off;
                if (p.mode == MODE_CONV && p.csplit > 0) off += (long long)(n0 / p.csplit) * p.st_split + (n0 % p.csplit);
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
                }  // !planar
                }  // row_ok
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    if (EPI_TMA && warp == 4 && lane == 0) bulk_wait_group<0>();   // smem staging tiles must outlive the reduce-stores
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
