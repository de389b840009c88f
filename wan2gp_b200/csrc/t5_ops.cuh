// Kernels of the umT5 text encoder (models/wan/modules/t5.py: T5Encoder with per-layer relative position embeddings), the step in front
// of the denoise path.  The encoder runs once per prompt on 512 tokens: its GEMMs (q|k|v, o, gate, fc1, fc2
// at M = 512) go through the tcgen05 pair GEMM; what is left -- embedding rows, RMS norm, a 64-wide attention with an additive
// position bias and a key mask, the gated-GELU product -- is small (4.3 GFLOP of attention per layer) and runs on the CUDA cores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100.cuh"

namespace b200 {

// out[i, :] = table[ids[i], :]  (token_embedding, t5.py:283); table bf16 or fp32, out fp32 (the residual stream stays fp32)
template <bool TABLE_BF16>
__global__ void embed_rows_kernel(const long long* __restrict__ ids, const void* __restrict__ table, float* __restrict__ out, int dim) {
    const long long id = ids[blockIdx.x];
    float* o = out + (long long)blockIdx.x * dim;
    if constexpr (TABLE_BF16) {
        const __nv_bfloat162* t = reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(table) + id * dim);
        for (int i = threadIdx.x; i < dim / 2; i += blockDim.x) {
            const float2 f = __bfloat1622float2(t[i]);
            o[2 * i] = f.x; o[2 * i + 1] = f.y;
        }
    } else {
        const float* t = reinterpret_cast<const float*>(table) + id * dim;
        for (int i = threadIdx.x; i < dim; i += blockDim.x) o[i] = t[i];
    }
}

// T5LayerNorm (t5.py:56-70): y = w * x * rsqrt(mean(x^2) + eps), fp32 statistics; bf16 output for the GEMM A operand, fp32 for the final norm
template <bool OUT_F32>
__global__ void __launch_bounds__(256) t5_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, void* __restrict__ out, int dim, float eps) {
    __shared__ float red[8];
    const float* xr = x + (long long)blockIdx.x * dim;
    float ss = 0.f;
    for (int i = threadIdx.x; i < dim; i += 256) { const float v = xr[i]; ss = fmaf(v, v, ss); }
    ss = block_sum_256(ss, red);
    const float r = rsqrtf(ss / (float)dim + eps);
    for (int i = threadIdx.x; i < dim; i += 256) {
        const float y = w[i] * (xr[i] * r);
        if constexpr (OUT_F32) reinterpret_cast<float*>(out)[(long long)blockIdx.x * dim + i] = y;
        else reinterpret_cast<__nv_bfloat16*>(out)[(long long)blockIdx.x * dim + i] = __float2bfloat16(y);
    }
}

// out = bf16(a * b): the gated FFN's fc1(x) * gelu(gate(x)) (t5.py:145), both factors bf16 GEMM outputs
__global__ void mul_bf16_kernel(const __nv_bfloat162* __restrict__ a, const __nv_bfloat162* __restrict__ b, __nv_bfloat162* __restrict__ out, long long n2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) {
        const float2 fa = __bfloat1622float2(a[i]), fb = __bfloat1622float2(b[i]);
        out[i] = __floats2bfloat162_rn(fa.x * fb.x, fa.y * fb.y);
    }
}

// T5Attention (t5.py:91-128) for one sequence, head dim 64: out = softmax(q k^T + bias[h, j - i] [key j >= n_valid -> masked]) v.
// No 1/sqrt(d) scaling.  bias_rel[h][j - i + L - 1] is the per-layer relative position embedding evaluated per offset (host side,
// wan/t5.py::relative_bias_table).  One block per (32 query rows, head): K^T and V of the head in shared memory (bf16), each warp walks
// 4 query rows: lanes own keys for the scores (K^T rows are contiguous over keys: conflict-free) and output dims for P V.
constexpr int T5_ATT_ROWS = 32, T5_ATT_D = 64;
__global__ void __launch_bounds__(256) t5_attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                           const __nv_bfloat16* __restrict__ v, long long ld, const float* __restrict__ bias_rel,
                                                           __nv_bfloat16* __restrict__ out, long long ldo, int L, int Lp, int n_valid) {
    extern __shared__ uint8_t t5_smem[];
    __nv_bfloat16* KT = reinterpret_cast<__nv_bfloat16*>(t5_smem);               // [64][Lp]
    __nv_bfloat16* V = KT + T5_ATT_D * Lp;                                       // [Lp][64]
    float* bias = reinterpret_cast<float*>(V + (long long)Lp * T5_ATT_D);        // [2 L - 1]
    float* P = bias + ((2 * L - 1 + 3) & ~3);                                    // [8 warps][Lp]
    const int h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = h * T5_ATT_D;
    for (int idx = threadIdx.x; idx < Lp * 32; idx += 256) {
        const int j = idx >> 5, dp = idx & 31;
        __nv_bfloat162 kk = __floats2bfloat162_rn(0.f, 0.f), vv = kk;
        if (j < L) {
            kk = *reinterpret_cast<const __nv_bfloat162*>(k + (long long)j * ld + col + 2 * dp);
            vv = *reinterpret_cast<const __nv_bfloat162*>(v + (long long)j * ld + col + 2 * dp);
        }
        KT[(2 * dp) * Lp + j] = kk.x;
        KT[(2 * dp + 1) * Lp + j] = kk.y;
        *reinterpret_cast<__nv_bfloat162*>(V + (long long)j * T5_ATT_D + 2 * dp) = vv;
    }
    for (int idx = threadIdx.x; idx < 2 * L - 1; idx += 256) bias[idx] = bias_rel[(long long)h * (2 * L - 1) + idx];
    __syncthreads();

    float* Pw = P + warp * Lp;
    const int nk = Lp / 32;                      // keys per lane (Lp <= 512 -> <= 16)
    for (int r = 0; r < T5_ATT_ROWS / 8; ++r) {
        const int i = blockIdx.x * T5_ATT_ROWS + warp * (T5_ATT_ROWS / 8) + r;
        if (i >= L) break;                       // warp-uniform
        float qf[T5_ATT_D];
        {
            const uint4* qp = reinterpret_cast<const uint4*>(q + (long long)i * ld + col);      // same address in every lane: one broadcast load
            #pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 u = qp[c];
                const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    qf[c * 8 + 2 * e] = __uint_as_float(wv[e] << 16);
                    qf[c * 8 + 2 * e + 1] = __uint_as_float(wv[e] & 0xffff0000u);
                }
            }
        }
        float s[16];
        float mx = -INFINITY;
        #pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            s[kk] = -INFINITY;
            if (kk < nk) {
                const int j = kk * 32 + lane;
                float acc = 0.f;
                #pragma unroll
                for (int d = 0; d < T5_ATT_D; ++d) acc = fmaf(qf[d], __bfloat162float(KT[d * Lp + j]), acc);
                if (j < n_valid) s[kk] = acc + bias[j - i + L - 1];          // keys >= n_valid: masked (t5.py:112-117); >= L: padding
                mx = fmaxf(mx, s[kk]);
            }
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        #pragma unroll
        for (int kk = 0; kk < 16; ++kk)
            if (kk < nk) { s[kk] = __expf(s[kk] - mx); sum += s[kk]; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        #pragma unroll
        for (int kk = 0; kk < 16; ++kk)                                      // softmax in fp32, probabilities rounded to bf16 (t5.py:121)
            if (kk < nk) Pw[kk * 32 + lane] = __bfloat162float(__float2bfloat16(s[kk] * inv));
        __syncwarp();
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < Lp; ++j) {
            const float pj = Pw[j];
            const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(V + (long long)j * T5_ATT_D + 2 * lane));
            a0 = fmaf(pj, vv.x, a0); a1 = fmaf(pj, vv.y, a1);
        }
        *reinterpret_cast<__nv_bfloat162*>(out + (long long)i * ldo + col + 2 * lane) = __floats2bfloat162_rn(a0, a1);
        __syncwarp();
    }
}

}  // namespace b200
