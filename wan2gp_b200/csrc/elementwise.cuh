// HBM-bound row kernels of the Wan DiT block: LayerNorm+modulation, full-width QK RMSNorm + RoPE,
// patch embed, head unpatchify, conditioning GEMVs, CFG + scheduler step.  One CTA per token row,
// 128-bit loads/stores, fp32 statistics, row kept in registers so every tensor is read once.
#pragma once
#include "sm100.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// y = LN(x) * (1 + scale) + shift            (WanLayerNorm no-affine + modulation, model.py:634-638, 686-691)
// y = LN(x) * w + b                          (norm3, model.py:664)         -> bf16
// x fp32 [L, D] (row stride D), mean/var two-pass in registers (matches torch layer_norm numerics).
constexpr int LN_MAXV = 8;    // float4 per thread, D <= 256*4*8 = 8192
__global__ void __launch_bounds__(256)
ln_modulate_kernel(const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
                   int scale_is_affine, int pre_round, __nv_bfloat16* __restrict__ y, int D, float eps) {
    __shared__ float red[8];
    const long long row = blockIdx.x;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const int nv = D >> 2;
    float4 v[LN_MAXV];
    float s = 0.f;
    #pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int idx = threadIdx.x + i * 256;
        if (idx < nv) { v[i] = __ldg(xr + idx); s += v[i].x + v[i].y + v[i].z + v[i].w; }
    }
    const float mean = block_sum_256(s, red) / D;
    float q = 0.f;
    #pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int idx = threadIdx.x + i * 256;
        if (idx < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += a * a + b * b + c * c + d * d;
        }
    }
    const float rstd = rsqrtf(block_sum_256(q, red) / D + eps);
    uint2* yr = reinterpret_cast<uint2*>(y + row * D);
    #pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int idx = threadIdx.x + i * 256;
        if (idx < nv) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + idx);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + idx);
            const float add = scale_is_affine ? 0.f : 1.f;
            float n0 = (v[i].x - mean) * rstd, n1 = (v[i].y - mean) * rstd, n2 = (v[i].z - mean) * rstd, n3 = (v[i].w - mean) * rstd;
            if (pre_round) {     // Hunyuan: LayerNorm output is cast to bf16 BEFORE the modulation (hyvideo/modules/models.py:211)
                n0 = __bfloat162float(__float2bfloat16_rn(n0)); n1 = __bfloat162float(__float2bfloat16_rn(n1));
                n2 = __bfloat162float(__float2bfloat16_rn(n2)); n3 = __bfloat162float(__float2bfloat16_rn(n3));
            }
            const float a = n0 * (add + sc.x) + sh.x;
            const float b = n1 * (add + sc.y) + sh.y;
            const float c = n2 * (add + sc.z) + sh.z;
            const float d = n3 * (add + sc.w) + sh.w;
            yr[idx] = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// In-place on a bf16 row of width D (all heads): x <- RoPE( x * rsqrt(mean(x^2)+eps) * w )
// WanRMSNorm over the FULL dim (model.py:152-175, production semantics) then the interleaved-pair rotation
// of posemb_layers.py:251-259 in fp32 with the [L,128] cos/sin tables; cos == nullptr skips RoPE
// (cross-attention q/k, model.py:255-258).
// PER_HEAD: statistics over each 128-wide head instead (Hunyuan RMSNorm over head_dim, hyvideo/modules/norm_layers.py:62-70;
// w is then [128]); the 16 lanes that hold one head reduce with shuffles.
constexpr int RN_MAXV = 4;    // uint4 (8 bf16) per thread, D <= 256*8*4 = 8192
template <bool PER_HEAD>
__global__ void __launch_bounds__(256)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ x2, long long ld, const float* __restrict__ w,
                    const float* __restrict__ w2, int D, float eps, const float* __restrict__ cos_t, const float* __restrict__ sin_t) {
    __shared__ float red[8];
    const long long row = blockIdx.x;
    // blockIdx.y = 1: the second segment (k of a fused q|k|v buffer) with its own norm weight -- q and k of one token in ONE launch,
    // their CTAs adjacent in the grid so the token's cos/sin row is fetched from HBM once
    if (blockIdx.y) { x = x2; w = w2; }
    uint4* xr = reinterpret_cast<uint4*>(x + row * ld);
    const int nv = D >> 3;
    uint4 v[RN_MAXV];
    float s = 0.f;
    float rh[RN_MAXV];
    #pragma unroll
    for (int i = 0; i < RN_MAXV; ++i) {
        const int idx = threadIdx.x + i * 256;
        float si = 0.f;
        if (idx < nv) {
            v[i] = xr[idx];
            const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = __uint_as_float(u[k] << 16), b = __uint_as_float(u[k] & 0xffff0000u);
                si += a * a + b * b;
            }
        }
        if (PER_HEAD) {      // 16 consecutive lanes == one head (nv is a multiple of 16, so a head never straddles the tail)
            #pragma unroll
            for (int o = 8; o > 0; o >>= 1) si += __shfl_xor_sync(0xffffffffu, si, o);
            rh[i] = rsqrtf(si / 128.f + eps);
        }
        s += si;
    }
    float r = 0.f;
    if (!PER_HEAD) r = rsqrtf(block_sum_256(s, red) / D + eps);
    // every element this thread owns sits at the same position d inside its head (the stride 256 * 8 is a multiple of 128): the
    // 16 table values are loaded ONCE per thread instead of once per 16-byte chunk (4 table LDG.128 per data LDG.128 made the
    // kernel load-issue bound at 45 % of HBM)
    float cv[8], sv[8];
    if (cos_t) {
        const int d = (threadIdx.x << 3) & 127;
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(cos_t + row * 128 + d));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(cos_t + row * 128 + d + 4));
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sin_t + row * 128 + d));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(sin_t + row * 128 + d + 4));
        cv[0] = c0.x; cv[1] = c0.y; cv[2] = c0.z; cv[3] = c0.w; cv[4] = c1.x; cv[5] = c1.y; cv[6] = c1.z; cv[7] = c1.w;
        sv[0] = s0.x; sv[1] = s0.y; sv[2] = s0.z; sv[3] = s0.w; sv[4] = s1.x; sv[5] = s1.y; sv[6] = s1.z; sv[7] = s1.w;
    }
    #pragma unroll
    for (int i = 0; i < RN_MAXV; ++i) {
        const int idx = threadIdx.x + i * 256;
        if (idx < nv) {
            const int col = idx << 3;
            if (PER_HEAD) r = rh[i];
            const int wcol = PER_HEAD ? (col & 127) : col;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + wcol));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + wcol + 4));
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            float f[8];
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                f[2 * k] = __uint_as_float(u[k] << 16) * r * wv[2 * k];
                f[2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u) * r * wv[2 * k + 1];
            }
            if (cos_t) {
                #pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float x0 = f[2 * k], x1 = f[2 * k + 1];
                    f[2 * k] = x0 * cv[2 * k] - x1 * sv[2 * k];
                    f[2 * k + 1] = x1 * cv[2 * k + 1] + x0 * sv[2 * k + 1];
                }
            }
            xr[idx] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same arithmetic (full-width RMSNorm + interleaved-pair RoPE, in place), software-pipelined.  ncu of the kernel above at the 14B
// shape (profiles/ncu_r02_rows.txt): 967 us for 3.19 GB = 3.3 TB/s, half of the HBM rate, with the long-scoreboard stall at 15 issue
// slots per instruction -- a CTA loads its 10 KB row, waits, reduces, stores, and only then does the next CTA's load start: at 4 CTAs per
// SM there are never more than 40 KB in flight per SM and nothing at all during the reduce / store phase.  Here a CTA walks RP_ITEMS
// consecutive (row, segment) items with a ring of RP_STAGES shared-memory rows filled by cp.async (LDGSTS): the loads of the next two
// rows are in flight while the current row is reduced, rotated and stored.  Every thread copies exactly the 16-byte chunks it later reads
// itself, so cp.async.wait_group is the only synchronisation the ring needs (no mbarrier, no __syncthreads beyond the block reduction).
constexpr int RP_STAGES = 3, RP_ITEMS = 8;
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256)
rmsnorm_rope_pipe_kernel(__nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ x2, long long ld, const float* __restrict__ w,
                         const float* __restrict__ w2, int D, float eps, const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                         long long n_items, int nseg) {
    extern __shared__ uint4 rp_smem[];                 // [RP_STAGES][nv]
    __shared__ float red[8];
    const int nv = D >> 3;
    const long long item0 = (long long)blockIdx.x * RP_ITEMS;
    const int n_my = (int)((n_items - item0) < RP_ITEMS ? (n_items - item0) : RP_ITEMS);
    auto row_ptr = [&](long long item) -> uint4* {
        const long long row = item / nseg;
        return reinterpret_cast<uint4*>(((item - row * nseg) ? x2 : x) + row * ld);
    };
    auto issue = [&](int it) {
        const uint4* g = row_ptr(item0 + it);
        uint4* sdst = rp_smem + (it % RP_STAGES) * nv;
        for (int idx = threadIdx.x; idx < nv; idx += 256) cp_async_16(sdst + idx, g + idx);
    };
    #pragma unroll
    for (int it = 0; it < RP_STAGES - 1; ++it) {
        if (it < n_my) issue(it);
        cp_async_commit();
    }
    for (int it = 0; it < n_my; ++it) {
        if (it + RP_STAGES - 1 < n_my) issue(it + RP_STAGES - 1);   // into the stage this thread finished reading one iteration ago
        cp_async_commit();
        cp_async_wait<RP_STAGES - 1>();                             // this thread's chunks of item `it` have landed
        const long long item = item0 + it;
        const long long row = item / nseg;
        const bool second = (item - row * nseg) != 0;
        const float* wr = second ? w2 : w;
        uint4* xr = row_ptr(item);
        const uint4* sr = rp_smem + (it % RP_STAGES) * nv;
        uint4 v[RN_MAXV];
        float s = 0.f;
        #pragma unroll
        for (int i = 0; i < RN_MAXV; ++i) {
            const int idx = threadIdx.x + i * 256;
            float si = 0.f;
            if (idx < nv) {
                v[i] = sr[idx];
                const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                #pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a = __uint_as_float(u[k] << 16), b = __uint_as_float(u[k] & 0xffff0000u);
                    si += a * a + b * b;
                }
            }
            s += si;
        }
        const float r = rsqrtf(block_sum_256(s, red) / D + eps);
        float cv[8], sv[8];
        if (cos_t) {
            const int d = (threadIdx.x << 3) & 127;
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(cos_t + row * 128 + d));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(cos_t + row * 128 + d + 4));
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(sin_t + row * 128 + d));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(sin_t + row * 128 + d + 4));
            cv[0] = c0.x; cv[1] = c0.y; cv[2] = c0.z; cv[3] = c0.w; cv[4] = c1.x; cv[5] = c1.y; cv[6] = c1.z; cv[7] = c1.w;
            sv[0] = s0.x; sv[1] = s0.y; sv[2] = s0.z; sv[3] = s0.w; sv[4] = s1.x; sv[5] = s1.y; sv[6] = s1.z; sv[7] = s1.w;
        }
        #pragma unroll
        for (int i = 0; i < RN_MAXV; ++i) {
            const int idx = threadIdx.x + i * 256;
            if (idx < nv) {
                const int col = idx << 3;
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr + col));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr + col + 4));
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                float f[8];
                #pragma unroll
                for (int k = 0; k < 4; ++k) {
                    f[2 * k] = __uint_as_float(u[k] << 16) * r * wv[2 * k];
                    f[2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u) * r * wv[2 * k + 1];
                }
                if (cos_t) {
                    #pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float x0 = f[2 * k], x1 = f[2 * k + 1];
                        f[2 * k] = x0 * cv[2 * k] - x1 * sv[2 * k];
                        f[2 * k + 1] = x1 * cv[2 * k + 1] + x0 * sv[2 * k + 1];
                    }
                }
                xr[idx] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            }
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (context / misc), n % 4 == 0
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

// ---------------------------------------------------------------------------------------------
// Patch embedding (Conv3d k=s=(1,2,2) == per-token GEMM, model.py:1131-1132, 1631, 1731) in fp32 (the
// reference locks this layer to fp32, model.py:1330-1371).  x = cat(latent[C0], y[C1]) channel-wise
// (model.py:1597-1600), [C,T,H,W] fp32; out [L, D] fp32, token order (t, h', w'), K order (c, ph, pw).
// Tile: 32 tokens x 128 channels per CTA of 256 threads; K walked in chunks of 32 through smem.
constexpr int PE_TOK = 32, PE_CH = 128, PE_KC = 32;
__global__ void __launch_bounds__(256)
patch_embed_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out, int T, int H, int W, int D, int P /* patch 1 or 2 */) {
    __shared__ float sx[PE_TOK][PE_KC + 1];
    __shared__ float sw[PE_CH][PE_KC + 1];
    const int K = (C0 + C1) * P * P;
    const int Hp = H / P, Wp = W / P;
    const int L = T * Hp * Wp;
    const int tok0 = blockIdx.x * PE_TOK, ch0 = blockIdx.y * PE_CH;
    // thread -> 4 tokens x 4 channels
    const int tc = threadIdx.x & 31, tt = threadIdx.x >> 5;      // channel lane (x4 strided by 32), token group (4 tokens)
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += PE_KC) {
        __syncthreads();
        for (int i = threadIdx.x; i < PE_TOK * PE_KC; i += 256) {
            const int tl = i / PE_KC, k = k0 + (i % PE_KC);
            const int l = tok0 + tl;
            float val = 0.f;
            if (l < L && k < K) {
                const int t = l / (Hp * Wp), r = l - t * (Hp * Wp), hp = r / Wp, wp = r - hp * Wp;
                const int c = k / (P * P), ph = (k / P) % P, pw = k % P;
                const float* src = c < C0 ? x0 + (long long)c * T * H * W : x1 + (long long)(c - C0) * T * H * W;
                val = __ldg(src + ((long long)t * H + (P * hp + ph)) * W + P * wp + pw);
            }
            sx[tl][i % PE_KC] = val;
        }
        for (int i = threadIdx.x; i < PE_CH * PE_KC; i += 256) {
            const int cl = i / PE_KC, k = k0 + (i % PE_KC);
            sw[cl][i % PE_KC] = (ch0 + cl < D && k < K) ? __ldg(w + (long long)(ch0 + cl) * K + k) : 0.f;
        }
        __syncthreads();
        #pragma unroll 8
        for (int k = 0; k < PE_KC; ++k) {
            float a[4], b[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sx[tt * 4 + i][k]; b[i] = sw[tc + 32 * i][k]; }
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int l = tok0 + tt * 4 + i;
        if (l >= L) continue;
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ch = ch0 + tc + 32 * j;
            if (ch < D) out[(long long)l * D + ch] = acc[i][j] + __ldg(bias + ch);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// unpatchify (model.py:2100-2126): y [L, 4*C] fp32 with feature order (ph, pw, c) -> out [C, T, H, W] fp32
// c_major = 1: feature order (c, ph, pw) (Hunyuan, hyvideo/modules/models.py:1235-1248); P = patch size 1 or 2
__global__ void unpatchify_kernel(const float* __restrict__ y, float* __restrict__ out, int C, int T, int H, int W, int P, int c_major) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)C * T * H * W;
    if (i >= n) return;
    const int w = i % W; long long r = i / W;
    const int h = r % H; r /= H;
    const int t = r % T; const int c = r / T;
    const int Hp = H / P, Wp = W / P;
    const long long l = ((long long)t * Hp + (h / P)) * Wp + (w / P);
    const int sub = (h % P) * P + (w % P);
    out[i] = __ldg(y + l * (P * P * C) + (c_major ? c * (P * P) + sub : sub * C + c));
}

// ---------------------------------------------------------------------------------------------
// out[n] = act_out( sum_k act_in(x[k]) * W[n,k] + b[n] )   fp32 GEMV, one warp per output (time embedding MLP,
// model.py:1141-1143, 1815-1818: Linear(256,D) SiLU Linear(D,D); time_projection = SiLU Linear(D,6D))
__device__ __forceinline__ float silu(float v) { return v / (1.f + __expf(-v)); }
__global__ void gemv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                float* __restrict__ out, int N, int K, int silu_in, int silu_out) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    const float4* wr = reinterpret_cast<const float4*>(w + (long long)n * K);
    const float4* xr = reinterpret_cast<const float4*>(x);
    for (int k = lane; k < (K >> 2); k += 32) {
        float4 xv = __ldg(xr + k);
        if (silu_in) { xv.x = silu(xv.x); xv.y = silu(xv.y); xv.z = silu(xv.z); xv.w = silu(xv.w); }
        const float4 wv = __ldg(wr + k);
        acc += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        acc += b[n];
        out[n] = silu_out ? silu(acc) : acc;
    }
}
// sinusoidal_embedding_1d (model.py:32-42): out[0:half] = cos(t * 10000^(-i/half)), out[half:] = sin(...)
// t_dev != null: the timestep is read from device memory (whole-step CUDA graph: the captured launch must not bake the value in)
__global__ void sinusoid_kernel(float t, const float* __restrict__ t_dev, float* __restrict__ out, int dim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim >> 1;
    if (i >= half) return;
    if (t_dev) t = __ldg(t_dev);
    const float a = t * powf(10000.f, -(float)i / (float)half);
    out[i] = cosf(a);
    out[half + i] = sinf(a);
}
// out[c] = mean over rows of x[rows, cols] (fp32): masked mean of the valid text tokens (hyvideo/modules/token_refiner.py:221-226)
__global__ void col_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += __ldg(x + (long long)r * cols + c);
    out[c] = s / rows;
}
// out[j] = a[j] + b[j]   (modulation tables: blocks.i.modulation + e0, head.modulation + e)
__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n, int bmod) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i % bmod];
}

// ---------------------------------------------------------------------------------------------
// CFG combine + flow-matching Euler update (any2video.py:1701-1722 plain CFG; euler_scheduler.py:67-86):
// lat <- lat - dt * (u + g (c - u)); also writes the combined prediction (optional)
// CFG-Zero* (any2video.py:1701-1722): alpha = <c,u> / (||u||^2 + 1e-8) over the whole sample, uncond *= alpha before the
// combine.  Bit-reproducible: every CTA writes its partial sums to scratch[2 + 2 b], the LAST CTA to finish (atomic ticket)
// adds the partials in index order in double and publishes dots[0] = sum c*u, dots[1] = sum u*u.  The same inputs therefore give
// the same alpha on every launch and on every rank of a CFG-pair split (float atomicAdd accumulation depended on CTA order).
// scratch: float[CFG_DOTS_FLOATS] = {dots[2], partials[2 * CFG_DOTS_MAX_BLOCKS], ticket}; the ticket must be 0 on entry and is
// reset on exit.
constexpr int CFG_DOTS_MAX_BLOCKS = 1184;
constexpr int CFG_DOTS_FLOATS = 2 + 2 * CFG_DOTS_MAX_BLOCKS + 2;
__global__ void __launch_bounds__(256)
cfg_dots_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, float* __restrict__ scratch, long long n4) {
    __shared__ float red[8];
    __shared__ bool last;
    float cu = 0.f, uu = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(cond) + i);
        const float4 u = __ldg(reinterpret_cast<const float4*>(uncond) + i);
        cu += c.x * u.x + c.y * u.y + c.z * u.z + c.w * u.w;
        uu += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
    }
    cu = block_sum_256(cu, red);
    uu = block_sum_256(uu, red);
    float* part = scratch + 2;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 2 + 2 * CFG_DOTS_MAX_BLOCKS);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = cu;
        part[2 * blockIdx.x + 1] = uu;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (unsigned int j = 0; j < gridDim.x; ++j) {
            a += (double)__ldcg(part + 2 * j);
            b += (double)__ldcg(part + 2 * j + 1);
        }
        scratch[0] = (float)a;
        scratch[1] = (float)b;
        *ticket = 0u;
    }
}

// gdt_dev != null: {guide, dt} are read from device memory (whole-step CUDA graph)
__global__ void cfg_euler_kernel(float* __restrict__ lat, const float* __restrict__ cond, const float* __restrict__ uncond,
                                 float g, float dt, const float* __restrict__ gdt_dev, float* __restrict__ pred_out,
                                 const float* __restrict__ star_dots, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    if (gdt_dev) { g = __ldg(gdt_dev); dt = __ldg(gdt_dev + 1); }
    const float4 c = __ldg(reinterpret_cast<const float4*>(cond) + i);
    float4 u = uncond ? __ldg(reinterpret_cast<const float4*>(uncond) + i) : c;
    if (star_dots) {
        const float alpha = __ldg(star_dots) / (__ldg(star_dots + 1) + 1e-8f);
        u.x *= alpha; u.y *= alpha; u.z *= alpha; u.w *= alpha;
    }
    float4 p = make_float4(u.x + g * (c.x - u.x), u.y + g * (c.y - u.y), u.z + g * (c.z - u.z), u.w + g * (c.w - u.w));
    float4 x = reinterpret_cast<float4*>(lat)[i];
    x.x -= dt * p.x; x.y -= dt * p.y; x.z -= dt * p.z; x.w -= dt * p.w;
    reinterpret_cast<float4*>(lat)[i] = x;
    if (pred_out) reinterpret_cast<float4*>(pred_out)[i] = p;
}

// One FlowUniPCMultistepScheduler.step (order <= 2, bh2, predict_x0, flow prediction) fused with the CFG combine; the scalar
// coefficients come from the host (pipeline.py::UniPCSchedule).  Everything fp32, one pass: 5 reads + 3 writes per element.
struct UniPCCoef { float sigma, ca, cb, cc, cd, pp, pq, pr; };
// pdev != null: {guide, sigma, ca, cb, cc, cd, pp, pq, pr, use_corrector} are read from device memory (whole-step CUDA graph)
__global__ void cfg_unipc_kernel(float* __restrict__ lat, const float* __restrict__ cond, const float* __restrict__ uncond, float g,
                                 float* __restrict__ x_last, const float* __restrict__ m0, float* __restrict__ m1, UniPCCoef k,
                                 int use_corrector, const float* __restrict__ pdev, const float* __restrict__ star_dots, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    if (pdev) {
        g = __ldg(pdev);
        k = UniPCCoef{__ldg(pdev + 1), __ldg(pdev + 2), __ldg(pdev + 3), __ldg(pdev + 4), __ldg(pdev + 5), __ldg(pdev + 6), __ldg(pdev + 7), __ldg(pdev + 8)};
        use_corrector = __ldg(pdev + 9) != 0.f;
    }
    const float4 c = __ldg(reinterpret_cast<const float4*>(cond) + i);
    float4 u = uncond ? __ldg(reinterpret_cast<const float4*>(uncond) + i) : c;
    if (star_dots) {
        const float alpha = __ldg(star_dots) / (__ldg(star_dots + 1) + 1e-8f);
        u.x *= alpha; u.y *= alpha; u.z *= alpha; u.w *= alpha;
    }
    const float v[4] = {u.x + g * (c.x - u.x), u.y + g * (c.y - u.y), u.z + g * (c.z - u.z), u.w + g * (c.w - u.w)};
    const float4 x4 = reinterpret_cast<const float4*>(lat)[i];
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(m0) + i);
    const float x[4] = {x4.x, x4.y, x4.z, x4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
    float xl[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (use_corrector) {
        const float4 l4 = reinterpret_cast<const float4*>(x_last)[i], b4 = reinterpret_cast<const float4*>(m1)[i];
        xl[0] = l4.x; xl[1] = l4.y; xl[2] = l4.z; xl[3] = l4.w;
        b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
    }
    float x0[4], xc[4], xn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        x0[j] = x[j] - k.sigma * v[j];
        xc[j] = use_corrector ? k.ca * xl[j] + k.cb * a[j] + k.cc * b[j] + k.cd * x0[j] : x[j];
        // pp = sigma_next / sigma_i is exactly 0 on the last step: skip the product so a non-finite xc cannot leak into x0
        xn[j] = (k.pp != 0.f ? k.pp * xc[j] : 0.f) + k.pq * x0[j] + (k.pr != 0.f ? k.pr * a[j] : 0.f);
    }
    reinterpret_cast<float4*>(lat)[i] = make_float4(xn[0], xn[1], xn[2], xn[3]);
    reinterpret_cast<float4*>(x_last)[i] = make_float4(xc[0], xc[1], xc[2], xc[3]);
    reinterpret_cast<float4*>(m1)[i] = make_float4(x0[0], x0[1], x0[2], x0[3]);
}

}  // namespace b200
