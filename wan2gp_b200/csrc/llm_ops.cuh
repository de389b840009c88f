// Kernels of the decoder-only LLM text encoders that sit in front of the Hunyuan denoise path: the language
// tower of Qwen2.5-VL-7B (Hunyuan Video 1.5) / llava-llama-3-8b (HunyuanVideo 1.0) as the reference runs them through transformers
// (models/hyvideo/text_encoder/text_encoder_1_5.py:86-117, 439-505; text_encoder/__init__.py): RMSNorm -> q|k|v -> rotate-half RoPE ->
// causal grouped-query attention (head dim 128) -> o; RMSNorm -> SwiGLU.  The encoder runs once per prompt on a few hundred tokens: its
// linear layers go through the tcgen05 GEMMs (bias / SiLU epilogues), the RMS norm, the embedding rows and the gated product reuse the T5
// row kernels (t5_ops.cuh); what is left -- RoPE in place on the fused q|k|v buffer and a causal attention of ~1e10 FLOP per layer -- runs
// on the CUDA cores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100.cuh"

namespace b200 {

constexpr int LLM_HD = 128;      // head dim of both towers

// rotate-half RoPE (transformers apply_rotary_pos_emb: x * cos + rotate_half(x) * sin, rotate_half(x) = cat(-x2, x1)) in place on the first
// `nheads` heads of every row of x (the q heads followed by the k heads of a fused q|k|v row); cos/sin fp32 [L, 64] (the table's two halves
// are equal: cat(freqs, freqs)).  Arithmetic in fp32, one bf16 rounding.
__global__ void __launch_bounds__(256) rope_half_kernel(__nv_bfloat16* __restrict__ x, long long ld, const float* __restrict__ cos_t,
                                                         const float* __restrict__ sin_t, int nheads) {
    __nv_bfloat16* xr = x + (long long)blockIdx.x * ld;
    const float* c = cos_t + (long long)blockIdx.x * (LLM_HD / 2);
    const float* s = sin_t + (long long)blockIdx.x * (LLM_HD / 2);
    for (int idx = threadIdx.x; idx < nheads * (LLM_HD / 2); idx += blockDim.x) {
        const int h = idx >> 6, i = idx & 63;
        const float a = __bfloat162float(xr[h * LLM_HD + i]), b = __bfloat162float(xr[h * LLM_HD + i + 64]);
        xr[h * LLM_HD + i] = __float2bfloat16(a * c[i] - b * s[i]);
        xr[h * LLM_HD + i + 64] = __float2bfloat16(b * c[i] + a * s[i]);
    }
}

// Causal grouped-query attention, head dim 128: out[i, h] = softmax_{j <= i}(scale * q[i, h] . k[j, h / group]) v[j, h / group].
// One warp per (query row, q head), 8 rows per block.  The query (pre-scaled by scale * log2 e) sits in shared memory; keys are walked in
// tiles of 32: lane l owns key j0 + l for the score (its own 256-byte K row, 16 x 128-bit loads) and output dims 4l..4l+3 for P V (one
// coalesced 256-byte V row per key, the probability broadcast by shuffle); online softmax across tiles (running maximum m, sum l).
// K / V of one kv head (L x 256 B each) are re-read by every row and every q head of the group: they live in L1 / L2.
__global__ void __launch_bounds__(256) causal_gqa_attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                                    const __nv_bfloat16* __restrict__ v, long long ldq, long long ldkv,
                                                                    __nv_bfloat16* __restrict__ out, long long ldo, int L, int group,
                                                                    float scale_log2e) {
    __shared__ float qs[8][LLM_HD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    const int h = blockIdx.y, hk = h / group;
    if (i >= L) return;                                             // warp-uniform; nothing below synchronises across warps
    {
        const uint2 u = *reinterpret_cast<const uint2*>(q + (long long)i * ldq + (long long)h * LLM_HD + 4 * lane);
        qs[warp][4 * lane + 0] = __uint_as_float(u.x << 16) * scale_log2e;
        qs[warp][4 * lane + 1] = __uint_as_float(u.x & 0xffff0000u) * scale_log2e;
        qs[warp][4 * lane + 2] = __uint_as_float(u.y << 16) * scale_log2e;
        qs[warp][4 * lane + 3] = __uint_as_float(u.y & 0xffff0000u) * scale_log2e;
    }
    __syncwarp();
    const __nv_bfloat16* kb = k + (long long)hk * LLM_HD;
    const __nv_bfloat16* vb = v + (long long)hk * LLM_HD;
    float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j0 = 0; j0 <= i; j0 += 32) {
        const int j = j0 + lane;
        float s = -INFINITY;
        if (j <= i) {
            const uint4* kr = reinterpret_cast<const uint4*>(kb + (long long)j * ldkv);
            float acc = 0.f;
            #pragma unroll
            for (int c = 0; c < LLM_HD / 8; ++c) {
                const uint4 u = kr[c];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc = fmaf(qs[warp][c * 8 + 2 * e], __uint_as_float(w[e] << 16), acc);
                    acc = fmaf(qs[warp][c * 8 + 2 * e + 1], __uint_as_float(w[e] & 0xffff0000u), acc);
                }
            }
            s = acc;
        }
        float tmax = s;                                             // lane 0 of every tile has j0 <= i: the tile maximum is finite
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        const float m_new = fmaxf(m, tmax);
        const float corr = exp2f(m - m_new);                        // first tile: exp2f(-inf) = 0
        const float p = (j <= i) ? exp2f(s - m_new) : 0.f;
        l = l * corr + warp_sum(p);
        a0 *= corr; a1 *= corr; a2 *= corr; a3 *= corr;
        const int nj = min(32, i - j0 + 1);                         // warp-uniform
        for (int jj = 0; jj < nj; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, p, jj);
            const uint2 u = *reinterpret_cast<const uint2*>(vb + (long long)(j0 + jj) * ldkv + 4 * lane);
            a0 = fmaf(pj, __uint_as_float(u.x << 16), a0);
            a1 = fmaf(pj, __uint_as_float(u.x & 0xffff0000u), a1);
            a2 = fmaf(pj, __uint_as_float(u.y << 16), a2);
            a3 = fmaf(pj, __uint_as_float(u.y & 0xffff0000u), a3);
        }
        m = m_new;
    }
    const float inv = 1.0f / l;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(out + (long long)i * ldo + (long long)h * LLM_HD + 4 * lane);
    o[0] = __floats2bfloat162_rn(a0 * inv, a1 * inv);
    o[1] = __floats2bfloat162_rn(a2 * inv, a3 * inv);
}

}  // namespace b200
