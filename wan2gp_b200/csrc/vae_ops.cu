// WanVAE decode kernels (channels-last bf16 activations) and their C ABI entry points.
// The convolutions reuse the tcgen05 GEMM main loop in MODE_CONV (gemm_sm100.cuh); everything else here is
// HBM-bound row work with 128-bit accesses.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "../../include/wan2gp_b200.h"
#include <stdlib.h>

#include "conv_sm100.cuh"
#include "conv2_sm100.cuh"
#include "gemm_sm100.cuh"
#include "host_util.h"

using namespace b200;

#define CHECK_LAUNCH(name)                                                                                   \
    do {                                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                                \
        if (e__ != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(e__)); \
        b200_count_launch();                                                                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------
// RMS_norm over channels (vae.py:85-103: F.normalize(x, dim=C) * sqrt(C) * gamma) + SiLU, per pixel.
// LPP lanes cooperate on one pixel, each holding up to NCH 16-byte chunks (C <= 8 * NCH * LPP).
// optional output geometry: write frames [t0, t0+Tc) of [T,H,W,C] replicate-padded as [Tc+pt, H+2ph, W+2pw, C]
struct RmsPad { int on, H, W, t0, pt, ph, pw; };
template <int LPP, int NCH>
__global__ void __launch_bounds__(256)
rms_silu_cl_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ y,
                   long long P, int C, int do_silu, RmsPad g) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long opix = gt / LPP;               // output pixel (of the padded slice when g.on)
    const int sub = (int)(gt % LPP);
    const int nchunk = C >> 3;
    const bool active = opix < P;
    long long pix = opix;                          // source pixel
    if (g.on && active) {
        const unsigned Ho = g.H + 2 * g.ph, Wo = g.W + 2 * g.pw;
        const unsigned op = (unsigned)opix;         // < 2^31 (host-checked): 32-bit divisions only
        const unsigned r = op / Wo;
        const int w = (int)(op - r * Wo), t = (int)(r / Ho), h = (int)(r - (unsigned)t * Ho);
        const int ts = max(g.t0 + t - g.pt, 0), hs = min(max(h - g.ph, 0), g.H - 1), ws = min(max(w - g.pw, 0), g.W - 1);
        pix = ((long long)ts * g.H + hs) * g.W + ws;
    }
    uint4 v[NCH];
    float ss = 0.f;
    #pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int ch = sub + i * LPP;
        if (active && ch < nchunk) {
            v[i] = __ldg(reinterpret_cast<const uint4*>(x + pix * C) + ch);
            const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = __uint_as_float(u[k] << 16), b = __uint_as_float(u[k] & 0xffff0000u);
                ss += a * a + b * b;
            }
        }
    }
    #pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = sqrtf((float)C) / fmaxf(sqrtf(ss), 1e-12f);
    #pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int ch = sub + i * LPP;
        if (active && ch < nchunk) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8 + 4));
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            float f[8];
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                f[2 * k] = __uint_as_float(u[k] << 16) * inv * g[2 * k];
                f[2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u) * inv * g[2 * k + 1];
            }
            if (do_silu) {
                #pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = silu_fast(f[k]);
            }
            reinterpret_cast<uint4*>(y + opix * C)[ch] =
                make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
    }
}

template <int LPP, int NCH = 2>
static int launch_rms(const void* x, const float* gamma, void* y, long long P, int C, int silu, cudaStream_t st, RmsPad g = RmsPad{}) {
    const long long threads = P * LPP;
    if ((threads + 255) / 256 > 0x7fffffffLL) return b200_set_error(B200_ERR_ARG, "rms_silu_cl: too many pixels for one launch");
    rms_silu_cl_kernel<LPP, NCH><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), gamma, reinterpret_cast<__nv_bfloat16*>(y), P, C, silu, g);
    CHECK_LAUNCH("rms_silu_cl");
    return B200_OK;
}

static int rms_dispatch(const void* x, const float* gamma, void* y, long long P, int C, int silu, cudaStream_t st, RmsPad g) {
    const int nchunk = C / 8;
    if (nchunk <= 4) return launch_rms<2>(x, gamma, y, P, C, silu, st, g);
    if (nchunk <= 8) return launch_rms<4>(x, gamma, y, P, C, silu, st, g);
    if (nchunk <= 16) return launch_rms<8>(x, gamma, y, P, C, silu, st, g);
    if (nchunk <= 32) return launch_rms<16>(x, gamma, y, P, C, silu, st, g);
    if (nchunk <= 64) return launch_rms<32>(x, gamma, y, P, C, silu, st, g);
    return launch_rms<32, 4>(x, gamma, y, P, C, silu, st, g);       // C <= 1024 (Hunyuan 1.5 VAE, 1024-channel levels)
}
extern "C" int b200_rms_silu_cl(const void* x, const float* gamma, void* y, long long P, int C, int silu, void* stream) {
    if (!x || !gamma || !y || P <= 0 || C % 8 || C > 1024) return b200_set_error(B200_ERR_ARG, "rms_silu_cl: bad argument (C=%d)", C);
    return rms_dispatch(x, gamma, y, P, C, silu, (cudaStream_t)stream, RmsPad{});
}
// RMS_norm -> SiLU -> F.pad(mode="replicate") of CausalConv3d (hunyuanvideo_15_vae.py:107-158, 217-250) in one pass, for the time
// slice [t0, t0+Tc): the normalised tensor is only ever written in the padded layout the conv reads.
extern "C" int b200_rms_silu_pad_cl(const void* x, const float* gamma, void* y, int T, int H, int W, int C, int silu, int t0, int Tc,
                                    int pt, int ph, int pw, void* stream) {
    if (!x || !gamma || !y || C % 8 || C > 1024 || T <= 0 || H <= 0 || W <= 0 || t0 < 0 || Tc <= 0 || t0 + Tc > T || pt < 0 || ph < 0 || pw < 0)
        return b200_set_error(B200_ERR_ARG, "rms_silu_pad_cl: bad argument (C=%d)", C);
    const long long P = (long long)(Tc + pt) * (H + 2 * ph) * (W + 2 * pw);
    if (P > 0x7fffffffLL) return b200_set_error(B200_ERR_ARG, "rms_silu_pad_cl: slice too large");
    return rms_dispatch(x, gamma, y, P, C, silu, (cudaStream_t)stream, RmsPad{1, H, W, t0, pt, ph, pw});
}

// ---------------------------------------------------------------------------------------------
// nearest-exact 2x upsample in space, channels-last (vae.py:105-111, 124-127)
__global__ void upsample2x_cl_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int T, int H, int W, int C8) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)T * 2 * H * 2 * W * C8;
    if (i >= n) return;
    const int c = i % C8; long long r = i / C8;
    const int wo = r % (2 * W); r /= (2 * W);
    const int ho = r % (2 * H); const int t = r / (2 * H);
    y[i] = __ldg(x + (((long long)t * H + (ho >> 1)) * W + (wo >> 1)) * C8 + c);
}
extern "C" int b200_upsample2x_cl(const void* x, void* y, int T, int H, int W, int C, void* stream) {
    if (!x || !y || C % 8) return b200_set_error(B200_ERR_ARG, "upsample2x_cl: bad argument");
    const long long n = (long long)T * 4 * H * W * (C / 8);
    upsample2x_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), T, H, W, C / 8);
    CHECK_LAUNCH("upsample2x_cl");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// z -> z*std + mean -> conv2 (1x1x1, 16x16, fp32) -> bf16 channels-last [T,H,W,16]   (vae.py:631-637)
__global__ void vae_prologue_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ stdv,
                                    const float* __restrict__ w, const float* __restrict__ b, __nv_bfloat16* __restrict__ out,
                                    long long npix) {
    __shared__ float sw[16 * 16], sb[16], sm[16], ss[16];
    if (threadIdx.x < 256) sw[threadIdx.x] = w[threadIdx.x];
    if (threadIdx.x < 16) { sb[threadIdx.x] = b[threadIdx.x]; sm[threadIdx.x] = mean[threadIdx.x]; ss[threadIdx.x] = stdv[threadIdx.x]; }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float x[16];
    #pragma unroll
    for (int c = 0; c < 16; ++c) x[c] = __ldg(z + c * npix + i) * ss[c] + sm[c];
    uint32_t o[8];
    #pragma unroll
    for (int co = 0; co < 16; co += 2) {
        float a0 = sb[co], a1 = sb[co + 1];
        #pragma unroll
        for (int c = 0; c < 16; ++c) { a0 = fmaf(sw[co * 16 + c], x[c], a0); a1 = fmaf(sw[(co + 1) * 16 + c], x[c], a1); }
        o[co >> 1] = pack_bf16x2(a0, a1);
    }
    uint4* op = reinterpret_cast<uint4*>(out + i * 16);
    op[0] = make_uint4(o[0], o[1], o[2], o[3]);
    op[1] = make_uint4(o[4], o[5], o[6], o[7]);
}
extern "C" int b200_vae_prologue(const float* z, const float* mean, const float* stdv, const float* w, const float* b, void* out,
                                 int T, int H, int W, void* stream) {
    if (!z || !mean || !stdv || !w || !b || !out) return b200_set_error(B200_ERR_ARG, "vae_prologue: null argument");
    const long long npix = (long long)T * H * W;
    vae_prologue_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        z, mean, stdv, w, b, reinterpret_cast<__nv_bfloat16*>(out), npix);
    CHECK_LAUNCH("vae_prologue");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// frames -> uint8 (vae.py:18-20)
__global__ void frames_to_u8_kernel(const float4* __restrict__ x, uchar4* __restrict__ out, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(x + i);
    auto cvt = [](float f) -> unsigned char {
        f = fminf(fmaxf(f, -1.f), 1.f);
        f = rintf((f + 1.f) * 127.5f);              // torch.round == round-half-to-even
        return (unsigned char)fminf(fmaxf(f, 0.f), 255.f);
    };
    out[i] = make_uchar4(cvt(v.x), cvt(v.y), cvt(v.z), cvt(v.w));
}
extern "C" int b200_frames_to_u8(const float* x, uint8_t* out, long long n, void* stream) {
    if (!x || !out || n <= 0 || n % 4) return b200_set_error(B200_ERR_ARG, "frames_to_u8: bad argument");
    const long long n4 = n / 4;
    frames_to_u8_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<uchar4*>(out), n4);
    CHECK_LAUNCH("frames_to_u8");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// Hunyuan VAE helpers (channels-last bf16)
// replicate padding: out [T+pt, H+2ph, W+2pw, C] (pt frames in FRONT), 16-byte chunks
__global__ void pad_replicate_cl_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int T, int H, int W, int C8, int pt, int ph, int pw) {
    // grid: x = 16-byte chunks of one padded row, y = padded row, z = padded frame
    const int Ho = H + 2 * ph, Wo = W + 2 * pw;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= Wo * C8) return;
    const int w = idx / C8, c = idx - w * C8;
    const int h = blockIdx.y, t = blockIdx.z;
    const int ts = max(t - pt, 0), hs = min(max(h - ph, 0), H - 1), ws = min(max(w - pw, 0), W - 1);
    y[(((long long)t * Ho + h) * Wo + w) * C8 + c] = __ldg(x + (((long long)ts * H + hs) * W + ws) * C8 + c);
}
extern "C" int b200_pad_replicate_cl(const void* x, void* y, int T, int H, int W, int C, int pt, int ph, int pw, void* stream) {
    if (!x || !y || C % 8 || T <= 0 || H <= 0 || W <= 0 || pt < 0 || ph < 0 || pw < 0 || H + 2 * ph > 65535 || T + pt > 65535)
        return b200_set_error(B200_ERR_ARG, "pad_replicate_cl: bad argument");
    const dim3 grid((unsigned)(((long long)(W + 2 * pw) * (C / 8) + 255) / 256), (unsigned)(H + 2 * ph), (unsigned)(T + pt));
    pad_replicate_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), T, H, W, C / 8, pt, ph, pw);
    CHECK_LAUNCH("pad_replicate_cl");
    return B200_OK;
}

// planar fp32 [C, P] -> channels-last bf16 [P, C*rep], channel c repeated rep times (z -> z.repeat_interleave(rep, dim=1),
// hunyuanvideo_15_vae.py:489-490; rep = 1 is a plain layout change)
__global__ void planar_to_cl_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, long long P, int rep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = P * C * rep;
    if (i >= n) return;
    const int co = i % (C * rep); const long long p = i / (C * rep);
    y[i] = __float2bfloat16_rn(__ldg(x + (long long)(co / rep) * P + p));
}
extern "C" int b200_planar_to_cl(const float* x, void* y, int C, long long P, int rep, void* stream) {
    if (!x || !y || C <= 0 || P <= 0 || rep <= 0) return b200_set_error(B200_ERR_ARG, "planar_to_cl: bad argument");
    const long long n = P * C * rep;
    planar_to_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), C, P, rep);
    CHECK_LAUNCH("planar_to_cl");
    return B200_OK;
}

// Hunyuan 1.5 VAE Upsample tail (hunyuanvideo_15_vae.py:309-338): h = conv output [T,H,W,F*Co] (F = 8 with temporal up-sampling,
// else 4), x = block input [T,H,W,Ci].  out [To, 2H, 2W, Co] = shuffle(h) + shuffle(repeat_interleave(x)), To = 2T-1 | T.
// The first frame of a temporal up-sample only unfolds in space: its channels are read as (r2 r3 c') with c' < 2Co and the
// first Co kept; the shortcut repeats by rep/2 there.
__global__ void hy_upsample_cl_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                      int T, int H, int W, int Ci, int Co, int temporal) {
    const int C8 = Co >> 3;
    // grid: x = 16-byte chunks of one output row (2W * C8), y = output row, z = output frame (32-bit index math only)
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= 2 * W * C8) return;
    const int wo = idx / C8, c0 = (idx - wo * C8) * 8;
    const int ho = blockIdx.y, to = blockIdx.z;
    const long long i = (((long long)to * 2 * H + ho) * 2 * W + wo) * C8 + (c0 >> 3);
    const int r2 = ho & 1, r3 = wo & 1, hh = ho >> 1, ww = wo >> 1;
    const int F = temporal ? 8 : 4;
    const int rep = F * Co / Ci;
    int f, hch, xbase, xrep;
    if (temporal && to == 0) { f = 0; hch = (r2 * 2 + r3) * (2 * Co); xbase = (r2 * 2 + r3) * (Ci / 4); xrep = rep / 2; }
    else if (temporal) { f = 1 + (to - 1) / 2; const int r1 = (to - 1) & 1; const int g = (r1 * 2 + r2) * 2 + r3; hch = g * Co; xbase = g * (Ci / 8); xrep = rep; }
    else { f = to; const int g = r2 * 2 + r3; hch = g * Co; xbase = g * (Ci / 4); xrep = rep; }
    const long long pix = ((long long)f * H + hh) * W + ww;
    const uint4 hv = __ldg(reinterpret_cast<const uint4*>(h + pix * (F * Co) + hch + c0));
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
    const __nv_bfloat16* xp = x + pix * Ci + xbase;
    float o[8];
    #pragma unroll
    for (int k = 0; k < 4; ++k) { o[2 * k] = __uint_as_float(hw[k] << 16); o[2 * k + 1] = __uint_as_float(hw[k] & 0xffff0000u); }
    #pragma unroll
    for (int k = 0; k < 8; ++k) o[k] += __bfloat162float(xp[(c0 + k) / xrep]);
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}
extern "C" int b200_hy_upsample_cl(const void* h, const void* x, void* out, int T, int H, int W, int Ci, int Co, int temporal, void* stream) {
    const int F = temporal ? 8 : 4;
    if (!h || !x || !out || Co % 8 || (F * Co) % Ci || (temporal && (F * Co / Ci) % 2)) return b200_set_error(B200_ERR_ARG, "hy_upsample_cl: bad argument");
    const int To = temporal ? 2 * T - 1 : T;
    if (T <= 0 || H <= 0 || W <= 0 || 2 * H > 65535 || To > 65535) return b200_set_error(B200_ERR_ARG, "hy_upsample_cl: bad extent");
    const dim3 grid((unsigned)((2LL * W * (Co / 8) + 255) / 256), (unsigned)(2 * H), (unsigned)To);
    hy_upsample_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(h), reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), T, H, W, Ci, Co, temporal);
    CHECK_LAUNCH("hy_upsample_cl");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused frame quantisation + all-gather over NVLink peer memory: every rank converts its fp32 frames to uint8 ONCE and
// stores the bytes straight into slot `rank` of the gather buffer of EVERY GPU of the box (peer-mapped pointers from the
// symmetric-memory rendezvous; NVSwitch gives each peer full bandwidth), instead of frames_to_u8 -> HBM -> ncclAllGather.
// A cross-GPU barrier (symmetric-memory signal pads) after the launch publishes the data.
struct PeerPtrs { uint8_t* p[16]; };
__global__ void frames_to_u8_allgather_kernel(const float4* __restrict__ x, PeerPtrs peers, int n_peers, long long slot_off, long long n16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    auto cvt = [](float f) -> uint32_t {
        f = fminf(fmaxf(f, -1.f), 1.f);
        f = rintf((f + 1.f) * 127.5f);
        return (uint32_t)fminf(fmaxf(f, 0.f), 255.f);
    };
    uint32_t w[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) {                      // 16 floats -> 16 bytes
        const float4 v = __ldg(x + 4 * i + k);
        w[k] = cvt(v.x) | (cvt(v.y) << 8) | (cvt(v.z) << 16) | (cvt(v.w) << 24);
    }
    const uint4 o = make_uint4(w[0], w[1], w[2], w[3]);
    for (int r = 0; r < n_peers; ++r) reinterpret_cast<uint4*>(peers.p[r] + slot_off)[i] = o;      // 16-B st.global on peer apertures
}
extern "C" int b200_frames_to_u8_allgather(const float* x, const uint64_t* peer_bufs, int n_peers, int rank, long long n,
                                           void* stream) {
    if (!x || !peer_bufs || n_peers <= 0 || n_peers > 16 || rank < 0 || rank >= n_peers || n <= 0 || n % 16)
        return b200_set_error(B200_ERR_ARG, "frames_to_u8_allgather: bad argument");
    PeerPtrs pp;
    for (int r = 0; r < 16; ++r) pp.p[r] = r < n_peers ? reinterpret_cast<uint8_t*>(peer_bufs[r]) : nullptr;
    const long long n16 = n / 16;
    frames_to_u8_allgather_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), pp, n_peers, (long long)rank * n, n16);
    CHECK_LAUNCH("frames_to_u8_allgather");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// Hunyuan 1.5 VAE encoder helpers (hunyuanvideo_15_vae.py Downsample :253-296, Encoder tail :423-430)
// Downsample tail: space(-time) -> channel shuffle of the conv output h [T,H,W,Ch] (Ch = Co/F) plus the group-mean shortcut of the
// same shuffle of x [T,H,W,Ci]  ->  out [To, H/2, W/2, Co];  F = 8 (temporal: To = 1 + (T-1)/2, first frame spatial-only and
// duplicated over the two channel halves) or 4.  One thread = 8 consecutive output channels of one output pixel.
__global__ void hy_downsample_cl_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                        int T, int H, int W, int Ci, int Co, int temporal, int ldh) {
    const int C8 = Co >> 3, Ho = H >> 1, Wo = W >> 1;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= Wo * C8) return;
    const int wo = idx / C8, o0 = (idx - wo * C8) * 8;
    const int ho = blockIdx.y, to = blockIdx.z;
    const int F = temporal ? 8 : 4;                                      // ldh: channel pitch of h (>= Co/F; conv outputs are padded to 16)
    const int Ch = Co / F;
    const bool first = temporal && to == 0;
    const int gs = first ? (F * Ci / Co) / 2 : F * Ci / Co;            // shortcut group size (:273, :279, :285)
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int oc = o0 + k;
        // conv branch: which (r1, r2, r3, c) of h lands in output channel oc
        int r, c;
        if (first) { const int o2 = oc % (4 * Ch); r = o2 / Ch; c = o2 - r * Ch; }       // cat([h_first, h_first]) (:270)
        else { r = oc / Ch; c = oc - r * Ch; }
        const int r1 = (temporal && !first) ? (r >> 2) : 0, r2 = (r >> 1) & 1, r3 = r & 1;
        const int tf = temporal ? (first ? 0 : 1 + 2 * (to - 1) + r1) : to;
        float v = __bfloat162float(h[(((long long)tf * H + 2 * ho + r2) * W + 2 * wo + r3) * ldh + c]);
        // shortcut: mean over gs consecutive channels of the shuffled x (channel j = r * Ci + c)
        float acc = 0.f;
        for (int g = 0; g < gs; ++g) {
            const int j = oc * gs + g;
            const int rr = j / Ci, cc = j - rr * Ci;
            const int q1 = (temporal && !first) ? (rr >> 2) : 0, q2 = (rr >> 1) & 1, q3 = rr & 1;
            const int tx = temporal ? (first ? 0 : 1 + 2 * (to - 1) + q1) : to;
            acc += __bfloat162float(x[(((long long)tx * H + 2 * ho + q2) * W + 2 * wo + q3) * Ci + cc]);
        }
        o[k] = v + acc / (float)gs;
    }
    reinterpret_cast<uint4*>(out)[(((long long)to * Ho + ho) * Wo + wo) * C8 + (o0 >> 3)] =
        make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}
extern "C" int b200_hy_downsample_cl(const void* h, int ldh, const void* x, void* out, int T, int H, int W, int Ci, int Co, int temporal,
                                     void* stream) {
    const int F = temporal ? 8 : 4;
    if (ldh < Co / F) return b200_set_error(B200_ERR_ARG, "hy_downsample_cl: ldh < Co / F");
    if (!h || !x || !out || T <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || Co % 8 || Co % F || (F * Ci) % Co || (temporal && ((T - 1) & 1)) ||
        (temporal && (F * Ci / Co) % 2))
        return b200_set_error(B200_ERR_ARG, "hy_downsample_cl: bad argument");
    const int To = temporal ? 1 + (T - 1) / 2 : T;
    if (H / 2 > 65535 || To > 65535) return b200_set_error(B200_ERR_ARG, "hy_downsample_cl: bad extent");
    const dim3 grid((unsigned)(((long long)(W / 2) * (Co / 8) + 255) / 256), (unsigned)(H / 2), (unsigned)To);
    hy_downsample_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(h), reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), T, H, W, Ci, Co, temporal, ldh);
    CHECK_LAUNCH("hy_downsample_cl");
    return B200_OK;
}
// y[p, c] = mean_g x[p, c*r + g]  (Encoder.forward shortcut "b (c r) f h w -> b c r f h w".mean(r), :424-425); bf16 [P,C] -> bf16 [P,C/r]
__global__ void group_mean_cl_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long P, int C, int r) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Co = C / r;
    if (i >= P * Co) return;
    const long long p = i / Co;
    const int c = (int)(i - p * Co);
    float acc = 0.f;
    for (int g = 0; g < r; ++g) acc += __bfloat162float(x[p * C + c * r + g]);
    y[i] = __float2bfloat16_rn(acc / (float)r);
}
extern "C" int b200_group_mean_cl(const void* x, void* y, long long P, int C, int r, void* stream) {
    if (!x || !y || P <= 0 || C <= 0 || r <= 0 || C % r) return b200_set_error(B200_ERR_ARG, "group_mean_cl: bad argument");
    const long long n = P * (C / r);
    group_mean_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), P, C, r);
    CHECK_LAUNCH("group_mean_cl");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// Tile seams of the tiled VAE decode / encode (vae.py:664-674 blend_v / blend_h): the first `ext` rows (columns) of tile b become a
// linear cross-fade from the last `ext` rows (columns) of its upper (left) neighbour a:  b = a (1 - k/ext) + b (k/ext).
// a [planes, ha, wa], b [planes, hb, wb] fp32 planar; vertical: wa == wb, horizontal: ha == hb.
__global__ void blend_edge_kernel(const float* __restrict__ a, float* __restrict__ b, long long planes, int ha, int wa, int hb, int wb,
                                  int ext, int vertical) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int rows = vertical ? ext : hb, cols = vertical ? wb : ext;
    if (i >= planes * rows * cols) return;
    const int x = (int)(i % cols), y = (int)((i / cols) % rows);
    const long long pl = i / ((long long)cols * rows);
    const int k = vertical ? y : x;
    const float wgt = (float)k / (float)ext;
    const float av = vertical ? a[(pl * ha + (ha - ext + y)) * wa + x] : a[(pl * ha + y) * wa + (wa - ext + x)];
    float* bp = b + (pl * hb + y) * wb + x;
    *bp = av * (1.f - wgt) + *bp * wgt;
}
extern "C" int b200_blend_edge_f32(const float* a, float* b, long long planes, int ha, int wa, int hb, int wb, int extent, int vertical,
                                   void* stream) {
    if (!a || !b || planes <= 0 || ha <= 0 || wa <= 0 || hb <= 0 || wb <= 0 || (vertical ? wa != wb : ha != hb))
        return b200_set_error(B200_ERR_ARG, "blend_edge_f32: bad argument");
    int ext = extent;
    if (vertical) { if (ext > ha) ext = ha; if (ext > hb) ext = hb; } else { if (ext > wa) ext = wa; if (ext > wb) ext = wb; }
    if (ext <= 0) return B200_OK;
    const long long n = planes * (vertical ? (long long)ext * wb : (long long)hb * ext);
    blend_edge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, planes, ha, wa, hb, wb, ext, vertical);
    CHECK_LAUNCH("blend_edge_f32");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// VAE encode helpers
// planar fp32 [C, P] -> channels-last bf16 [P, Cpad], channels >= C zero (video [3,T,H,W] -> 8-channel TMA-legal operand)
__global__ void planar_to_cl_pad_kernel(const float* __restrict__ x, uint4* __restrict__ y, int C, long long P, int Cpad8) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * Cpad8) return;
    const long long p = i / Cpad8;
    const int c0 = (int)(i - p * Cpad8) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < C) ? __ldg(x + (long long)(c0 + j) * P + p) : 0.f;
    y[i] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
extern "C" int b200_planar_to_cl_pad(const float* x, void* y, int C, long long P, int Cpad, void* stream) {
    if (!x || !y || C <= 0 || P <= 0 || Cpad < C || Cpad % 8) return b200_set_error(B200_ERR_ARG, "planar_to_cl_pad: bad argument");
    const long long n = P * (Cpad / 8);
    planar_to_cl_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<uint4*>(y), C, P, Cpad / 8);
    CHECK_LAUNCH("planar_to_cl_pad");
    return B200_OK;
}
// space-to-depth 2x2: x [T,H,W,C] -> y [T,H/2,W/2,4C], channel (2p+q)*C + c = x[t, 2i+p, 2j+q, c]  (H, W even).  A stride-2
// 3x3 conv with ZeroPad2d((0,1,0,1)) (Resample 'downsample2d/3d', vae.py:134-143) becomes a stride-1 2x2 conv over y.
__global__ void space_to_depth_cl_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int C8) {
    // grid: x = 16-byte chunks of one OUTPUT row (W/2 * 4 * C8), y = output row, z = frame
    const int Wo = W >> 1, n = Wo * 4 * C8;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    const int j = idx / (4 * C8), r = idx - j * 4 * C8;
    const int pq = r / C8, c = r - pq * C8;
    const int i = blockIdx.y, t = blockIdx.z;
    y[(((long long)t * (H >> 1) + i) * Wo) * 4 * C8 + idx] =
        __ldg(x + (((long long)t * H + 2 * i + (pq >> 1)) * W + 2 * j + (pq & 1)) * C8 + c);
}
extern "C" int b200_space_to_depth_cl(const void* x, void* y, int T, int H, int W, int C, void* stream) {
    if (!x || !y || T <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C % 8 || H / 2 > 65535 || T > 65535)
        return b200_set_error(B200_ERR_ARG, "space_to_depth_cl: bad argument (H, W even; C %% 8 == 0)");
    const dim3 grid((unsigned)(((long long)(W / 2) * 4 * (C / 8) + 255) / 256), (unsigned)(H / 2), (unsigned)T);
    space_to_depth_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), H, W, C / 8);
    CHECK_LAUNCH("space_to_depth_cl");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// HunyuanVideo 1.0 VAE helpers: GroupNorm over the WHOLE clip (torch.nn.GroupNorm on [B,C,T,H,W], unet_causal_3d_blocks.py
// :378/:399, vae.py:292) on channels-last bf16.  Pass 1: per-group sum / sum-of-squares (fp32 per thread and block, fp64
// across blocks, fixed summation order).  Pass 2 (apply) normalises, optionally applies SiLU and writes straight into the replicate-PADDED layout
// the following causal conv reads, for a time slice [t0, t0+Tc) of the clip -- the un-padded normalised tensor never exists.
__global__ void __launch_bounds__(256)
group_stats_kernel(const uint4* __restrict__ x, float* __restrict__ part, long long P, int C8, int Cg, int G) {
    __shared__ float sh[2][256][8];
    const int tid = threadIdx.x;
    const int rows = 256 / C8;                      // C8 divides 256 (host-checked)
    const int c8 = tid % C8, r = tid / C8;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    for (long long p = (long long)blockIdx.x * rows + r; p < P; p += (long long)gridDim.x * rows) {
        const uint4 v = __ldg(x + p * C8 + c8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h[j]);
            s[2 * j] += f.x; q[2 * j] += f.x * f.x;
            s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { sh[0][tid][j] = s[j]; sh[1][tid][j] = q[j]; }
    __syncthreads();
    // fixed-order reduction (no atomics: the statistics, hence the decode, are bit-reproducible run to run)
    if (tid < G) {
        float a = 0.f, b = 0.f;
        const int c_lo = tid * Cg, c_hi = c_lo + Cg;
        for (int rr = 0; rr < rows; ++rr)
            for (int c = c_lo; c < c_hi; ++c) { a += sh[0][rr * C8 + (c >> 3)][c & 7]; b += sh[1][rr * C8 + (c >> 3)][c & 7]; }
        part[((long long)blockIdx.x * G + tid) * 2] = a;
        part[((long long)blockIdx.x * G + tid) * 2 + 1] = b;
    }
}
// per-channel affine of the normalisation: y = x * scale[c] + shift[c], scale = rstd_g * gamma[c], shift = beta[c] - mean_g * scale
__global__ void group_stats_finalize_kernel(const float* __restrict__ part, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            float* __restrict__ scale_shift, int nblocks, double inv_n, float eps, int G, int Cg) {
    __shared__ float s_mean[256], s_rstd[256];
    const int g = threadIdx.x;
    if (g < G) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < nblocks; ++i) { a += (double)part[((long long)i * G + g) * 2]; b += (double)part[((long long)i * G + g) * 2 + 1]; }
        const double mean = a * inv_n;
        const double var = fmax(b * inv_n - mean * mean, 0.0);    // biased variance, as torch group_norm
        s_mean[g] = (float)mean;
        s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int C = G * Cg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = s_rstd[c / Cg] * gamma[c];
        scale_shift[c] = sc;
        scale_shift[C + c] = beta[c] - s_mean[c / Cg] * sc;
    }
}
#define GROUP_STATS_MAX_BLOCKS (148 * 8)
extern "C" int b200_group_stats_cl(const void* x, const float* gamma, const float* beta, float* stats, void* workspace, long long P, int C,
                                   int G, float eps, void* stream) {
    if (!x || !gamma || !beta || !stats || !workspace || P <= 0 || G <= 0 || G > 256 || C % 8 || C % G) return b200_set_error(B200_ERR_ARG, "group_stats_cl: bad argument");
    const int C8 = C / 8, Cg = C / G;
    if (C8 > 256 || 256 % C8) return b200_set_error(B200_ERR_ARG, "group_stats_cl: C/8 = %d must divide 256", C8);
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = 256 / C8;
    const long long want = (P + rows - 1) / rows;
    const unsigned grid = (unsigned)(want < GROUP_STATS_MAX_BLOCKS ? want : GROUP_STATS_MAX_BLOCKS);
    group_stats_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<float*>(workspace), P, C8, Cg, G);
    CHECK_LAUNCH("group_stats_cl");
    group_stats_finalize_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const float*>(workspace), gamma, beta, stats, (int)grid,
                                                   1.0 / ((double)P * Cg), eps, G, Cg);
    CHECK_LAUNCH("group_stats_finalize");
    return B200_OK;
}

constexpr int GN_CPT = 4;          // 16-byte chunks per thread, all loaded before any arithmetic (bytes in flight)
__global__ void __launch_bounds__(256)
group_norm_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ scale_shift, uint4* __restrict__ y, int H, int W, int C8,
                        int c8_shift, int silu, int t0, int pt, int ph, int pw) {
    // grid: x = 16-byte chunks of one padded output row (Wo * C8) / (256 * GN_CPT), y = padded row h, z = padded frame t
    const int Wo = W + 2 * pw, Ho = H + 2 * ph, n = Wo * C8, C = C8 * 8;
    const int h = blockIdx.y, t = blockIdx.z;
    const int ts = max(t0 + t - pt, 0), hs = min(max(h - ph, 0), H - 1);
    const uint4* xrow = x + ((long long)ts * H + hs) * W * C8;
    uint4* yrow = y + ((long long)t * Ho + h) * Wo * C8;
    uint4 v[GN_CPT];
    int c8s[GN_CPT];
    #pragma unroll
    for (int k = 0; k < GN_CPT; ++k) {
        const int idx = (blockIdx.x * GN_CPT + k) * 256 + threadIdx.x;
        const int w = c8_shift >= 0 ? (idx >> c8_shift) : idx / C8;
        c8s[k] = idx - w * C8;
        const int ws = min(max(w - pw, 0), W - 1);
        v[k] = idx < n ? __ldg(xrow + (long long)ws * C8 + c8s[k]) : make_uint4(0, 0, 0, 0);
    }
    #pragma unroll
    for (int k = 0; k < GN_CPT; ++k) {
        const int idx = (blockIdx.x * GN_CPT + k) * 256 + threadIdx.x;
        if (idx >= n) continue;
        const float4* sc = reinterpret_cast<const float4*>(scale_shift + c8s[k] * 8);
        const float4* sh = reinterpret_cast<const float4*>(scale_shift + C + c8s[k] * 8);
        const float4 a0 = __ldg(sc), a1 = __ldg(sc + 1), b0 = __ldg(sh), b1 = __ldg(sh + 1);
        const float A[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float B[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint32_t u[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        float f[8];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[2 * j] = fmaf(__uint_as_float(u[j] << 16), A[2 * j], B[2 * j]);
            f[2 * j + 1] = fmaf(__uint_as_float(u[j] & 0xffff0000u), A[2 * j + 1], B[2 * j + 1]);
        }
        if (silu) {
            #pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = silu_fast(f[j]);
        }
        yrow[idx] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
}
extern "C" int b200_group_norm_apply_cl(const void* x, const float* scale_shift, void* y, int T, int H, int W, int C, int silu, int t0,
                                        int Tc, int pt, int ph, int pw, void* stream) {
    if (!x || !scale_shift || !y || C % 8 || T <= 0 || H <= 0 || W <= 0 || t0 < 0 || Tc <= 0 || t0 + Tc > T || pt < 0 || ph < 0 || pw < 0)
        return b200_set_error(B200_ERR_ARG, "group_norm_apply_cl: bad argument");
    if (H + 2 * ph > 65535 || Tc + pt > 65535) return b200_set_error(B200_ERR_ARG, "group_norm_apply_cl: slice too large");
    const int C8 = C / 8;
    int shift = -1;                                   // C/8 a power of two: index split by shift instead of a division
    if ((C8 & (C8 - 1)) == 0) { shift = 0; while ((1 << shift) < C8) ++shift; }
    const long long chunks = (long long)(W + 2 * pw) * C8;
    const dim3 grid((unsigned)((chunks + 256 * GN_CPT - 1) / (256 * GN_CPT)), (unsigned)(H + 2 * ph), (unsigned)(Tc + pt));
    group_norm_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x), scale_shift, reinterpret_cast<uint4*>(y), H, W, C8, shift, silu, t0, pt, ph, pw);
    CHECK_LAUNCH("group_norm_apply_cl");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// causal conv as implicit GEMM
// optional explicit view for conv_cl_impl: input extents + tap origin offsets, output strides (elements) -- used by the
// phase-decomposed up-sampling convs of the HunyuanVideo 1.0 VAE
struct ConvView { int Ti, Hi, Wi, off_t, off_h, off_w; long long ost_t, ost_h, ost_w; };
static int conv_cl_impl(const void* x, const void* w, const float* bias, const void* residual, void* out, int T, int H, int W,
                        int Cin, int Cout, int kt, int kh, int kw, int out_mode, int t_off, int pad_h, int pad_w, int up_py, int up_px,
                        void* stream, int prepadded = 0, const ConvView* view = nullptr, const float* norm_gamma = nullptr,
                        void* norm_out = nullptr, int t_valid = 0);

extern "C" int b200_conv3d_cl(const void* x, const void* w, const float* bias, const void* residual, void* out, int T, int H,
                              int W, int Cin, int Cout, int kt, int kh, int kw, int out_mode, int t_off, void* stream) {
    return conv_cl_impl(x, w, bias, residual, out, T, H, W, Cin, Cout, kt, kh, kw, out_mode, t_off, kh >> 1, kw >> 1, -1, -1, stream);
}

// nearest-exact 2x upsample followed by a 3x3 conv (vae.py:124-133) == four 2x2 convs on the LOW-resolution input, one per
// output parity (py, px), with pre-summed taps (w4: [4 phases][Cout][4 taps][Cin], phase = 2*py+px, taps (a,b) row-major):
// 2.25x fewer FLOPs and the 4x larger up-sampled tensor is never written.  out: bf16 [T, 2H, 2W, Cout].
extern "C" int b200_upconv2x_cl(const void* x, const void* w4, const float* bias, void* out, int T, int H, int W, int Cin, int Cout,
                                void* stream) {
    if (!x || !w4 || !out) return b200_set_error(B200_ERR_ARG, "upconv2x_cl: null argument");
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(w4) + (long long)(2 * py + px) * Cout * 4 * Cin;
            // parity 0 reads source rows (h-1, h); parity 1 reads (h, h+1)
            int r = conv_cl_impl(x, wp, bias, nullptr, out, T, H, W, Cin, Cout, 1, 2, 2, 0, 0, 1 - py, 1 - px, py, px, stream);
            if (r) return r;
        }
    return B200_OK;
}

static bool conv_row_wanted(int W, int kh, int kw);
static bool conv_norm_instance(int BN, int Cout);
static int env_flag(const char* name, int dflt);
// 1 if b200_conv3d_cl_norm / b200_upconv2x_cl_norm can serve a layer with these dimensions (row-tiled kernel, one N tile of 96 or 192
// channels).  W is the INPUT width of the launch (the low-resolution width for the up-sampling conv).
extern "C" int b200_conv_norm_fusable(int W, int Cin, int Cout, int kh, int kw) {
    const int BN = b200_pick_bn(Cout, false);
    if (!conv_row_wanted(W, kh, kw) || !conv_norm_instance(BN, Cout) || Cin % 8) return 0;
    static const int row_k32n = env_flag("B200_CONV_ROW_K32", 1);
    const bool k96 = (Cin == 96) && row_k32n;
    return (BN == 192 && k96) ? 0 : 1;
}

// b200_conv3d_cl (out_mode 0) whose epilogue ALSO writes silu(RMS_norm(out) * gamma) -- the input of the next layer's conv -- to
// norm_out (bf16 [T,H,W,Cout]); out may be NULL when the raw tensor has no other consumer.  Replaces CausalConv3d followed by
// RMS_norm + SiLU of the next residual sub-layer (vae.py:246-250 inside ResidualBlock.forward :254-273).
extern "C" int b200_conv3d_cl_norm(const void* x, const void* w, const float* bias, const void* residual, void* out, void* norm_out,
                                   const float* gamma, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw, void* stream) {
    if (!norm_out || !gamma) return b200_set_error(B200_ERR_ARG, "conv3d_cl_norm: norm_out / gamma required");
    return conv_cl_impl(x, w, bias, residual, out, T, H, W, Cin, Cout, kt, kh, kw, 0, 0, kh >> 1, kw >> 1, -1, -1, stream, 0, nullptr, gamma, norm_out);
}
// Streaming (time-sliced) form of b200_conv3d_cl / b200_conv3d_cl_norm: x is [T + kt - 1, H, W, Cin] = the kt-1 history frames of the
// previous time slice (zeros for the first slice) followed by the T frames of this slice; the conv is 'valid' in time (no causal zero
// fill) and zero-padded in space; T output frames.  This is the reference's chunked decode with per-conv feature caches
// (vae.py:639-655, CausalConv3d.forward :55-61 cache_x) for slices of any length.  out_mode 0 (bf16 [T,H,W,Cout], optional residual and
// fused norm_out / gamma as b200_conv3d_cl_norm) or 1 (time_conv interleave, t_off as b200_conv3d_cl).
extern "C" int b200_conv3d_cl_stream(const void* x_hist, const void* w, const float* bias, const void* residual, void* out, void* norm_out,
                                     const float* gamma, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw, int out_mode, int t_off,
                                     void* stream) {
    if (out_mode != 0 && out_mode != 1) return b200_set_error(B200_ERR_ARG, "conv3d_cl_stream: out_mode 0 or 1");
    if (norm_out && out_mode != 0) return b200_set_error(B200_ERR_ARG, "conv3d_cl_stream: fused norm needs out_mode 0");
    return conv_cl_impl(x_hist, w, bias, residual, out, T, H, W, Cin, Cout, kt, kh, kw, out_mode, t_off, kh >> 1, kw >> 1, -1, -1, stream, 0, nullptr,
                        norm_out ? gamma : nullptr, norm_out, 1);
}

// b200_upconv2x_cl with the same fused norm: out and norm_out are [T, 2H, 2W, Cout].
extern "C" int b200_upconv2x_cl_norm(const void* x, const void* w4, const float* bias, void* out, void* norm_out, const float* gamma, int T,
                                     int H, int W, int Cin, int Cout, void* stream) {
    if (!x || !w4 || !out || !norm_out || !gamma) return b200_set_error(B200_ERR_ARG, "upconv2x_cl_norm: null argument");
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(w4) + (long long)(2 * py + px) * Cout * 4 * Cin;
            int r = conv_cl_impl(x, wp, bias, nullptr, out, T, H, W, Cin, Cout, 1, 2, 2, 0, 0, 1 - py, 1 - px, py, px, stream, 0, nullptr, gamma, norm_out);
            if (r) return r;
        }
    return B200_OK;
}

// Replicate-padded causal conv (Hunyuan VAEs, hunyuanvideo_15_vae.py:124-158): TMA zero fill cannot replicate, so the
// caller materialises the padded tensor once (b200_pad_replicate_cl) and this runs a "valid" conv over it.
// xpad bf16 [T+kt-1, H+kh-1, W+kw-1, Cin]; T,H,W are the OUTPUT dims.
extern "C" int b200_conv3d_cl_prepadded(const void* xpad, const void* w, const float* bias, const void* residual, void* out, int T,
                                        int H, int W, int Cin, int Cout, int kt, int kh, int kw, int out_mode, void* stream) {
    return conv_cl_impl(xpad, w, bias, residual, out, T, H, W, Cin, Cout, kt, kh, kw, out_mode, 0, 0, 0, -1, -1, stream, 1);
}

// "valid" conv over an arbitrary window of a channels-last tensor x [Ti,Hi,Wi,Cin]: output pixel (t,h,w), t<T, h<H, w<W, reads
// taps at x[off_t + t + dt, off_h + h + dh, off_w + w + dw] and is stored (bf16) at out + t*ost_t + h*ost_h + w*ost_w (elements).
// One launch per phase of a nearest-up-sample + conv pair (UpsampleCausal3D, unet_causal_3d_blocks.py:196-222): the conv
// over the 2x (x2x2) up-sampled tensor equals 4 (8) small convs with pre-summed taps over the LOW-resolution tensor.
extern "C" int b200_conv3d_cl_view(const void* x, int Ti, int Hi, int Wi, int off_t, int off_h, int off_w, const void* w, const float* bias,
                                   const void* residual, void* out, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw,
                                   long long ost_t, long long ost_h, long long ost_w, void* stream) {
    // taps that fall beyond the high end of the input read zeros (TMA out-of-bounds fill): that is the ZeroPad2d((0,1,0,1)) of the
    // encoder's stride-2 convs; the window must START inside the input
    if (off_t < 0 || off_h < 0 || off_w < 0 || off_t >= Ti || off_h >= Hi || off_w >= Wi)
        return b200_set_error(B200_ERR_ARG, "conv3d_cl_view: window origin outside the input");
    ConvView v{Ti, Hi, Wi, off_t, off_h, off_w, ost_t, ost_h, ost_w};
    return conv_cl_impl(x, w, bias, residual, out, T, H, W, Cin, Cout, kt, kh, kw, 0, 0, 0, 0, -1, -1, stream, 1, &v);
}

// ---- row-tiled conv kernel (conv_sm100.cuh): instances and selection
template <int BN, int ROWS, int BKC = 64, bool NORM = false, bool PAIR = false>
static int launch_conv_row_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    auto kern = conv_row_tcgen05_kernel<BN, ROWS, BKC, NORM, PAIR>;
    using S = ConvRowSmem<BN, ROWS, BKC, PAIR>;
    static std::atomic<unsigned long long> attr_done{0};
    if (b200_first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "conv_row smem attr: %s", cudaGetErrorString(e));
        b200_mark_used_on_device(attr_done);
    }
    if constexpr (PAIR) {
        // one cluster of two CTAs per pair of pixel tiles; persistent over min(#pairs, #SMs / 2) clusters
        const int pairs = (p.m_tiles + 1) / 2;
        const int clusters = pairs < b200_num_sms() / 2 ? pairs : b200_num_sms() / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = S::kBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "conv_row pair launch: %s", cudaGetErrorString(e));
        CHECK_LAUNCH("conv_row_pair_tcgen05");
        return B200_OK;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < b200_num_sms() ? tiles : b200_num_sms();
    kern<<<grid, 256, S::kBytes, st>>>(ta, tb, p);
    CHECK_LAUNCH("conv_row_tcgen05");
    return B200_OK;
}
// ---- second-generation pair kernel (conv2_sm100.cuh): unrolled taps, grouped weight stages, two epilogue warpgroups
template <int BN, int ROWS, int BKC, bool NORM, int KHW, int GROUP>
static int launch_conv_row2_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    auto kern = conv_row2_tcgen05_kernel<BN, ROWS, BKC, NORM, KHW, GROUP>;
    using S = ConvRow2Smem<BN, ROWS, BKC, KHW, GROUP>;
    static std::atomic<unsigned long long> attr_done{0};
    if (b200_first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "conv_row2 smem attr: %s", cudaGetErrorString(e));
        b200_mark_used_on_device(attr_done);
    }
    const int pairs = (p.m_tiles + 1) / 2;
    const int clusters = pairs < b200_num_sms() / 2 ? pairs : b200_num_sms() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(CONV2_THREADS);
    cfg.dynamicSmemBytes = S::kBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "conv_row2 launch: %s", cudaGetErrorString(e));
    CHECK_LAUNCH("conv_row2_tcgen05");
    return B200_OK;
}
// B200_CONV_V2=0 keeps every pair conv on the first-generation kernel (A/B measurements).  -> 1: no instance for this layer
static int launch_conv_row2(int BN, bool k32, bool norm, int khw, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    static const int on = env_flag("B200_CONV_V2", 1);
    if (!on) return 1;
#define B200_CONV2(BN_, ROWS_, BKC_, KHW_, G_)                                                                   \
    if (BN == BN_ && k32 == (BKC_ == 32) && khw == KHW_)                                                         \
        return norm ? launch_conv_row2_inst<BN_, ROWS_, BKC_, true, KHW_, G_>(ta, tb, p, st)                      \
                    : launch_conv_row2_inst<BN_, ROWS_, BKC_, false, KHW_, G_>(ta, tb, p, st);
    B200_CONV2(96, 2, 32, 3, 9)        // Wan decoder, full-resolution stage (Cin = 96 as 3 x 32 channels): one weight stage per (dt, chunk)
    B200_CONV2(96, 2, 64, 2, 4)        // folded 2x up-sampling 192 -> 96 (four 2x2 parity convs)
    B200_CONV2(192, 1, 64, 3, 3)       // 192-channel stage: one kh row of taps per weight stage
    B200_CONV2(192, 1, 64, 2, 4)       // folded 2x up-sampling 384 -> 192
#undef B200_CONV2
    if (!norm && !k32 && khw == 3) {   // Hunyuan VAE decoders (GroupNorm: no fused norm)
        if (BN == 128) return launch_conv_row2_inst<128, 2, 64, false, 3, 3>(ta, tb, p, st);
        if (BN == 256) return launch_conv_row2_inst<256, 1, 64, false, 3, 3>(ta, tb, p, st);
    }
    return 1;
}
static int conv_row_rows(int BN) { return BN <= 128 ? 2 : 1; }
// instances with the fused next-layer norm epilogue: the single-N-tile layers of the Wan decoder's 96- and 192-channel stages
static bool conv_norm_instance(int BN, int Cout) { return Cout == BN && (BN == 96 || BN == 192); }
static int launch_conv_row_norm(int BN, bool k32, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    if (pair) {
        if (BN == 96) return k32 ? launch_conv_row_inst<96, 2, 32, true, true>(ta, tb, p, st) : launch_conv_row_inst<96, 2, 64, true, true>(ta, tb, p, st);
        if (BN == 192 && !k32) return launch_conv_row_inst<192, 1, 64, true, true>(ta, tb, p, st);
    } else {
        if (BN == 96) return k32 ? launch_conv_row_inst<96, 2, 32, true>(ta, tb, p, st) : launch_conv_row_inst<96, 2, 64, true>(ta, tb, p, st);
        if (BN == 192 && !k32) return launch_conv_row_inst<192, 1, 64, true>(ta, tb, p, st);
    }
    return b200_set_error(B200_ERR_ARG, "no fused-norm row-conv instance for BN=%d", BN);
}
// CTA-pair instances (weights shared by two pixel tiles): the shared-memory-bound widths of the decoders' high-resolution stages
static bool conv_pair_instance(int BN, bool k32) { return (BN == 96) || (!k32 && (BN == 192 || BN == 128 || BN == 256)); }
static int launch_conv_row_pair(int BN, bool k32, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    if (BN == 96) return k32 ? launch_conv_row_inst<96, 2, 32, false, true>(ta, tb, p, st) : launch_conv_row_inst<96, 2, 64, false, true>(ta, tb, p, st);
    if (!k32) {
        switch (BN) {
            case 256: return launch_conv_row_inst<256, 1, 64, false, true>(ta, tb, p, st);
            case 192: return launch_conv_row_inst<192, 1, 64, false, true>(ta, tb, p, st);
            case 128: return launch_conv_row_inst<128, 2, 64, false, true>(ta, tb, p, st);
        }
    }
    return b200_set_error(B200_ERR_ARG, "no pair row-conv instance for BN=%d", BN);
}
static int launch_conv_row(int BN, bool k32, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    if (k32) {                                    // Cin = 96: three 32-channel chunks (64B swizzle)
        switch (BN) {
            case 96: return launch_conv_row_inst<96, 2, 32>(ta, tb, p, st);
            case 16: return launch_conv_row_inst<16, 2, 32>(ta, tb, p, st);
        }
        return b200_set_error(B200_ERR_ARG, "no K=32 row-conv instance for BN=%d", BN);
    }
    switch (BN) {
        case 256: return launch_conv_row_inst<256, 1>(ta, tb, p, st);
        case 192: return launch_conv_row_inst<192, 1>(ta, tb, p, st);
        case 128: return launch_conv_row_inst<128, 2>(ta, tb, p, st);
        case 96: return launch_conv_row_inst<96, 2>(ta, tb, p, st);
        case 64: return launch_conv_row_inst<64, 2>(ta, tb, p, st);
        case 32: return launch_conv_row_inst<32, 2>(ta, tb, p, st);
        case 16: return launch_conv_row_inst<16, 2>(ta, tb, p, st);
    }
    return b200_set_error(B200_ERR_ARG, "no row-conv instance for BN=%d", BN);
}
// B200_CONV_ROW=0 forces the per-tap kernel, =2 the row kernel for every spatial conv (A/B measurements).
// B200_CONV_ROW_BASEOFF=1 sets the descriptors' base-offset field to (addr >> 7) & 7: MEASURED WRONG on B200 -- the swizzle of
// both TMA and tcgen05.mma is a function of the absolute shared-memory address, so a start address that is not 1024-B
// aligned needs base offset 0 (round-1 A/B run: parity passes with 0, fails with the computed offset).
static int env_flag(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
// the row tile is 128 pixels wide: use it when at most ~20% of the last tile is padding and there are spatial taps to share
static bool conv_row_wanted(int W, int kh, int kw) {
    static const int on = env_flag("B200_CONV_ROW", 1);
    if (!on || kh * kw <= 1) return false;
    const int tiles = (W + CONVR_BW - 1) / CONVR_BW;
    return on == 2 || (double)W >= 0.8 * tiles * CONVR_BW;
}

static int conv_cl_impl(const void* x, const void* w, const float* bias, const void* residual, void* out, int T, int H, int W,
                        int Cin, int Cout, int kt, int kh, int kw, int out_mode, int t_off, int pad_h, int pad_w, int up_py, int up_px,
                        void* stream, int prepadded, const ConvView* view, const float* norm_gamma, void* norm_out, int t_valid) {
    if (!x || !w || (!out && !norm_out) || T <= 0 || H <= 0 || W <= 0) return b200_set_error(B200_ERR_ARG, "conv3d_cl: null/empty argument");
    if (Cin % 8) return b200_set_error(B200_ERR_ARG, "conv3d_cl: Cin %% 8 != 0");
    if (out_mode != 2 && Cout % 16) return b200_set_error(B200_ERR_ARG, "conv3d_cl: Cout %% 16 != 0");
    if (out_mode == 1 && (Cout % 64 || residual)) return b200_set_error(B200_ERR_ARG, "conv3d_cl: bad interleave arguments");
    const int taps = kt * kh * kw;
    const int BN = b200_pick_bn(out_mode == 2 ? 16 : Cout, false);
    const bool row = conv_row_wanted(W, kh, kw);
    const int ROWS = conv_row_rows(BN);
    // Cin = 96 (the full-resolution stage): K per tap is walked as 3 x 32 channels with 64B-swizzled boxes instead of
    // 2 x 64 with a half-empty second box (25% fewer MMAs and smem bytes)
    static const int row_k32 = env_flag("B200_CONV_ROW_K32", 1);
    const bool k96 = (Cin == 96) && (BN == 96 || BN == 16 || (BN == 32 && !row)) && (!row || row_k32);
    const uint32_t kbox = k96 ? 32 : 64;
    CUtensorMap ta, tb;
    {
        // t_valid: the input carries its kt-1 history frames explicitly ([T+kt-1,H,W,Cin], streaming decode) -- 'valid' in time, padded in space
        uint64_t Ti = (prepadded || t_valid) ? T + kt - 1 : T, Hi = prepadded ? H + kh - 1 : H, Wi = prepadded ? W + kw - 1 : W;
        if (view) { Ti = view->Ti; Hi = view->Hi; Wi = view->Wi; }
        uint64_t dims[4] = {(uint64_t)Cin, Wi, Hi, Ti};
        uint64_t str[3] = {(uint64_t)Cin * 2, Wi * Cin * 2, Hi * Wi * Cin * 2};
        uint32_t box[4] = {kbox, CONV_BW, CONV_BH, 1};
        if (row) { box[1] = CONVR_BW + kw - 1; box[2] = ROWS + kh - 1; }      // halo tile shared by all kh x kw taps
        int r = b200_make_tmap_bf16(&ta, x, 4, dims, str, box, k96 ? 64 : 128);
        if (r) return r;
    }
    // CTA pairs sharing the weight tile (B200_CONV_PAIR=0: single-CTA kernels everywhere, for A/B measurements): one N tile, row kernel,
    // enough tiles to fill the chip; outputs through the plain strided mapping (no interleave / planar modes)
    static const int pair_on = env_flag("B200_CONV_PAIR", 1);
    const bool pair = pair_on && row && Cout == BN && conv_pair_instance(BN, k96) && out_mode != 1 && out_mode != 2 &&
                      (long long)T * ((H + ROWS - 1) / ROWS) * ((W + CONVR_BW - 1) / CONVR_BW) >= 2LL * b200_num_sms();
    {
        uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)taps, (uint64_t)Cout};
        uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)taps * Cin * 2};
        uint32_t box[3] = {kbox, 1, (uint32_t)(pair ? BN / 2 : BN)};
        int r = b200_make_tmap_bf16(&tb, w, 3, dims, str, box, k96 ? 64 : 128);
        if (r) return r;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.mode = MODE_CONV;
    p.N = Cout;
    p.T = T; p.H = H; p.W = W; p.kt = kt; p.kh = kh; p.kw = kw;
    p.pad_h = pad_h; p.pad_w = pad_w; p.pad_t = (prepadded || t_valid) ? 0 : kt - 1;
    p.cin_chunks = k96 ? (row ? 3 : 1) : (Cin + 63) / 64;      // row kernel: three 32-channel chunks; per-tap k96: one 3-box stage
    p.num_k_iters = taps * p.cin_chunks;
    p.tiles_h = row ? (H + ROWS - 1) / ROWS : (H + CONV_BH - 1) / CONV_BH;
    p.tiles_w = row ? (W + CONVR_BW - 1) / CONVR_BW : (W + CONV_BW - 1) / CONV_BW;
    p.m_tiles = T * p.tiles_h * p.tiles_w;
    p.M = p.m_tiles * GEMM_BM;
    p.n_tiles = (Cout + BN - 1) / BN;
    p.n_group = p.n_tiles;            // weights are small: all N tiles of one pixel tile run back to back
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    if (view) {
        if (out_mode != 0) return b200_set_error(B200_ERR_ARG, "conv3d_cl_view: bf16 output only");
        p.pad_t = -view->off_t; p.pad_h = -view->off_h; p.pad_w = -view->off_w;
        p.out = out;
        p.st_t = view->ost_t; p.st_h = view->ost_h; p.st_w = view->ost_w;
    } else if (out_mode == 0 && up_py >= 0) {
        // output pixel (t, 2h + py, 2w + px) of a [T, 2H, 2W, Cout] tensor
        p.out = reinterpret_cast<__nv_bfloat16*>(out) + ((long long)up_py * 2 * W + up_px) * Cout;
        p.st_w = 2LL * Cout; p.st_h = 4LL * W * Cout; p.st_t = 4LL * H * W * Cout;
    } else if (out_mode == 0) {
        p.out = out;
        p.st_w = Cout; p.st_h = (long long)W * Cout; p.st_t = (long long)H * W * Cout;
    } else if (out_mode == 1) {
        const int C = Cout / 2;
        p.out = reinterpret_cast<__nv_bfloat16*>(out) + (long long)t_off * H * W * C;
        p.st_w = C; p.st_h = (long long)W * C; p.st_t = 2LL * H * W * C;
        p.csplit = C; p.st_split = (long long)H * W * C;
    } else if (out_mode == 3) {
        p.out = out; p.out_fp32 = 1;                                       // fp32 channels-last [T,H,W,Cout] (+ bf16 residual)
        p.st_w = Cout; p.st_h = (long long)W * Cout; p.st_t = (long long)H * W * Cout;
    } else if (out_mode == 2) {
        p.out = out; p.out_fp32 = 1; p.planar = 1;
        p.st_w = 1; p.st_h = W; p.st_t = (long long)H * W; p.st_split = (long long)T * H * W;
    } else {
        return b200_set_error(B200_ERR_ARG, "conv3d_cl: out_mode %d", out_mode);
    }
    if (norm_out) {
        // fused RMS_norm + SiLU of the consuming layer: bf16 channels-last outputs of the row kernel with all channels in one N tile
        if (!row || !conv_norm_instance(BN, Cout) || !norm_gamma || (out_mode != 0) || p.csplit || view)
            return b200_set_error(B200_ERR_ARG, "conv3d_cl_norm: layer not fusable (W=%d Cout=%d mode=%d): ask b200_conv_norm_fusable first", W, Cout, out_mode);
        p.norm_gamma = norm_gamma; p.norm_out = norm_out; p.norm_only = out ? 0 : 1;
        if (up_py >= 0) p.norm_out = reinterpret_cast<__nv_bfloat16*>(norm_out) + ((long long)up_py * 2 * W + up_px) * Cout;
        if (!out) p.out = p.norm_out;
        static const int base_off_n = env_flag("B200_CONV_ROW_BASEOFF", 0);
        p.conv_base_offset = base_off_n;
        if (pair && kh == kw) {
            const int r2 = launch_conv_row2(BN, k96, true, kh, ta, tb, p, (cudaStream_t)stream);
            if (r2 != 1) return r2;
        }
        return launch_conv_row_norm(BN, k96, pair, ta, tb, p, (cudaStream_t)stream);
    }
    if (row) {
        static const int base_off = env_flag("B200_CONV_ROW_BASEOFF", 0);
        p.conv_base_offset = base_off;
        if (pair && kh == kw && !p.csplit && !p.planar) {
            const int r2 = launch_conv_row2(BN, k96, false, kh, ta, tb, p, (cudaStream_t)stream);
            if (r2 != 1) return r2;
        }
        if (pair) return launch_conv_row_pair(BN, k96, ta, tb, p, (cudaStream_t)stream);
        return launch_conv_row(BN, k96, ta, tb, p, (cudaStream_t)stream);
    }
    return k96 ? b200_launch_gemm_k96(BN, ta, tb, p, (cudaStream_t)stream) : b200_launch_gemm(BN, false, ta, tb, p, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Decoder head conv (96|128 -> 3 channels, 3x3x3, planar fp32 frames): vae.py:505-508 `head`, Hunyuan `conv_out`.
// With Cout = 3 an implicit GEMM whose N is the output channels wastes the tensor core (N padded to 16) and -- what actually bounds
// it -- re-reads the A tile from shared memory once per tap: 162 M128xN16xK16 MMAs per 128 pixels, 37 ms at 720p x 81 frames = 6 % of
// HBM speed (VERDICT r01 "weak" #5).  Here the 9 spatial taps are STACKED INTO N instead:
//     G[t, h', w', (dh, dw, co)] = sum_{dt, ci} w[co, dt, dh, dw, ci] * x[t - 2 + dt, h', w', ci]          (a 3x1x1 conv, Cout' = 27 -> 32)
//     out[co, t, h, w]           = bias[co] + sum_{dh, dw} G[t, h + dh - 1, w + dw - 1, (dh, dw, co)]     (gather of 27 floats per pixel)
// Same multiply-adds (each product appears once), 18 MMAs of N = 32 per 128 pixels instead of 162 of N = 16, every activation byte
// goes through shared memory 3 times (frame taps) instead of 27; G (fp32, 128 B per pixel) makes one round trip through HBM/L2.
// Zero-padded convs (Wan): neighbours outside the frame contribute nothing.  Replicate-padded (Hunyuan): x is the pre-padded tensor
// [T+2, H+2, W+2, C], G covers its H+2 x W+2 pixels and every neighbour exists.
__global__ void __launch_bounds__(256)
head_gather_kernel(const float* __restrict__ G, const float* __restrict__ bias, float* __restrict__ out, int T, int H, int W, int Hg, int Wg,
                   int off, int Cout) {
    const int w = blockIdx.x * 256 + threadIdx.x;
    const int h = blockIdx.y, t = blockIdx.z;
    if (w >= W) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    #pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
        const int hh = h + dh - off;
        if (hh < 0 || hh >= Hg) continue;
        #pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
            const int ww = w + dw - off;
            if (ww < 0 || ww >= Wg) continue;
            const float* g = G + (((long long)t * Hg + hh) * Wg + ww) * 32 + (dh * 3 + dw) * Cout;
            for (int c = 0; c < Cout; ++c) acc[c] += __ldg(g + c);
        }
    }
    const long long plane = (long long)T * H * W;
    const long long pix = ((long long)t * H + h) * W + w;
    for (int c = 0; c < Cout; ++c) out[c * plane + pix] = acc[c] + __ldg(bias + c);
}

// x: bf16 channels-last [T,H,W,Cin] (prepadded = 0, zero padding) or the replicate-padded [T+2,H+2,W+2,Cin] (prepadded = 1);
// w_stack: bf16 [32][3 (dt)][Cin], row (dh*3+dw)*Cout + co, rows >= 9*Cout zero; bias fp32 [Cout]; ws: fp32 scratch of
// T*Hg*Wg*32 floats (Hg = H or H+2); out: planar fp32 [Cout, T, H, W].  Cout <= 3.
extern "C" int b200_conv3d_head_cl(const void* x, const void* w_stack, const float* bias, void* ws, long long ws_bytes, float* out, int T, int H,
                                   int W, int Cin, int Cout, int prepadded, void* stream) {
    if (!x || !w_stack || !bias || !ws || !out || Cout < 1 || Cout > 3 || T <= 0 || H <= 0 || W <= 0 || H > 65535 || T > 65535)
        return b200_set_error(B200_ERR_ARG, "conv3d_head_cl: bad argument");
    // prepadded: 0 = zero padding everywhere (whole clip); 1 = replicate-padded input [T+2,H+2,W+2]; 2 = streaming slice: 2 history frames
    // in front ([T+2,H,W]), zero padding in space
    const bool rep = prepadded == 1;
    const int Hg = rep ? H + 2 : H, Wg = rep ? W + 2 : W;
    if (ws_bytes < (long long)T * Hg * Wg * 32 * 4) return b200_set_error(B200_ERR_ARG, "conv3d_head_cl: workspace too small");
    // stacked 3x1x1 conv -> fp32 channels-last G (out_mode 3); 'valid' in time when the input carries its own front frames
    int r = conv_cl_impl(x, w_stack, nullptr, nullptr, ws, T, Hg, Wg, Cin, 32, 3, 1, 1, 3, 0, 0, 0, -1, -1, stream, rep ? 1 : 0, nullptr, nullptr, nullptr,
                         prepadded == 2 ? 1 : 0);
    if (r) return r;
    const dim3 grid((unsigned)((W + 255) / 256), (unsigned)H, (unsigned)T);
    head_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(ws), bias, out, T, H, W, Hg, Wg, rep ? 0 : 1, Cout);
    CHECK_LAUNCH("head_gather");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// single-head attention for the VAE middle block (1% of decode FLOPs): S = q k^T (tcgen05 GEMM, fp32),
// row softmax (this kernel), O = P v (tcgen05 GEMM with v as an MN-major B operand).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int N, long long lds, long long ldp, float scale_log2) {
    extern __shared__ float row[];
    __shared__ float red[8];
    // N = real keys; N8 = columns of the score row that exist (N rounded up to 8): the tail gets probability 0
    const int N8 = (N + 7) & ~7;
    const float* sr = s + (long long)blockIdx.x * lds;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < N; i += 256) { const float v = sr[i] * scale_log2; row[i] = v; mx = fmaxf(mx, v); }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    #pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    float sum = 0.f;
    for (int i = threadIdx.x; i < N; i += 256) { const float e = exp2f(row[i] - mx); row[i] = e; sum += e; }
    const float inv = 1.f / block_sum_256(sum, red);
    __nv_bfloat16* pr = p + (long long)blockIdx.x * ldp;
    for (int i = threadIdx.x; i < N8; i += 256) pr[i] = __float2bfloat16_rn(i < N ? row[i] * inv : 0.f);
}

// same for rows that do not fit the shared-memory row buffer: TWO passes over the global row -- pass 1 keeps an online
// (max, sum) pair per thread (flash-style rescale), pass 2 writes the probabilities; 16-byte loads, 8-byte stores.
__global__ void __launch_bounds__(256)
softmax_rows_long_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int N, long long lds, long long ldp, float scale_log2) {
    __shared__ float red[8];
    __shared__ float red_m[8];
    const int N8 = (N + 7) & ~7;
    const int N4 = N & ~3;
    const float* sr = s + (long long)blockIdx.x * lds;
    const float4* sr4 = reinterpret_cast<const float4*>(sr);
    float mx = -INFINITY, sum = 0.f;
    for (int i = threadIdx.x; i < N4 / 4; i += 256) {
        const float4 v = sr4[i];
        const float a = v.x * scale_log2, b = v.y * scale_log2, c = v.z * scale_log2, d = v.w * scale_log2;
        const float m4 = fmaxf(fmaxf(a, b), fmaxf(c, d));
        if (m4 > mx) { sum *= exp2f(mx - m4); mx = m4; }
        sum += exp2f(a - mx) + exp2f(b - mx) + exp2f(c - mx) + exp2f(d - mx);
    }
    for (int i = N4 + threadIdx.x; i < N; i += 256) {
        const float a = sr[i] * scale_log2;
        if (a > mx) { sum *= exp2f(mx - a); mx = a; }
        sum += exp2f(a - mx);
    }
    // block combine: global max, then rescaled sums
    float gm = mx;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    if ((threadIdx.x & 31) == 0) red_m[threadIdx.x >> 5] = gm;
    __syncthreads();
    gm = red_m[0];
    #pragma unroll
    for (int i = 1; i < 8; ++i) gm = fmaxf(gm, red_m[i]);
    sum = (mx == -INFINITY) ? 0.f : sum * exp2f(mx - gm);
    const float inv = 1.f / block_sum_256(sum, red);
    __nv_bfloat16* pr = p + (long long)blockIdx.x * ldp;
    uint2* pr4 = reinterpret_cast<uint2*>(pr);
    for (int i = threadIdx.x; i < N4 / 4; i += 256) {
        const float4 v = sr4[i];
        pr4[i] = make_uint2(pack_bf16x2(exp2f(v.x * scale_log2 - gm) * inv, exp2f(v.y * scale_log2 - gm) * inv),
                            pack_bf16x2(exp2f(v.z * scale_log2 - gm) * inv, exp2f(v.w * scale_log2 - gm) * inv));
    }
    for (int i = N4 + threadIdx.x; i < N8; i += 256) pr[i] = __float2bfloat16_rn(i < N ? exp2f(sr[i] * scale_log2 - gm) * inv : 0.f);
}

extern "C" int b200_attention_1head(const void* qkv, void* out, void* workspace, long long workspace_bytes, int F, int N, int C,
                                    float scale, int causal_frames, void* stream) {
    if (!qkv || !out || !workspace) return b200_set_error(B200_ERR_ARG, "attention_1head: null argument");
    if (C % 64) return b200_set_error(B200_ERR_ARG, "attention_1head: C %% 64 != 0");
    const int N8 = (N + 7) & ~7;      // GEMM extents are multiples of 8: the key tail [N, N8) gets probability 0 (the caller
                                      // guarantees 8 readable rows after the last frame of qkv)
    // causal_frames: frame f attends to the tokens of frames 0..f (Hunyuan 1.5 VAE mid block, hunyuanvideo_15_vae.py:161-214);
    // 2: full attention over all F*N tokens (HunyuanVideo 1.0 VAE with mid_block_causal_attn off); 0: every frame attends to
    // itself only (Wan VAE).
    const long long Lk_max = causal_frames ? (long long)F * N : N;
    const long long Np = (Lk_max + 63) / 64 * 64;     // padded row pitch so P is a legal GEMM operand
    const long long need = (long long)N * Np * 4 + (long long)N * Np * 2;
    (void)N8;
    if (workspace_bytes < need) return b200_set_error(B200_ERR_ARG, "attention_1head: workspace %lld < %lld bytes", workspace_bytes, need);
    float* S = reinterpret_cast<float*>(workspace);
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(S + (long long)N * Np);
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(qkv);
    cudaStream_t st = (cudaStream_t)stream;
    static std::atomic<unsigned long long> attr_done{0};
    if (b200_first_use_on_device(attr_done)) {
        cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        b200_mark_used_on_device(attr_done);
    }

    for (int f = 0; f < F; ++f) {
        const __nv_bfloat16* q = base + (long long)f * N * 3 * C;
        const __nv_bfloat16* kv = causal_frames ? base : q;                 // keys/values start at frame 0 when causal
        const int Lk = causal_frames == 1 ? (f + 1) * N : causal_frames == 2 ? F * N : N;   // 2: every frame sees all frames
        const int Lk8 = (Lk + 7) & ~7;
        int r = b200_gemm_bf16(q, kv + C, S, N, Lk8, C, 3LL * C, 3LL * C, Np, nullptr, nullptr, nullptr, 0, 1, 0, 0, stream);
        if (r) return r;
        if ((long long)Lk * 4 <= 200 * 1024) softmax_rows_kernel<<<N, 256, Lk * sizeof(float), st>>>(S, P, Lk, Np, Np, scale * 1.4426950408889634f);
        else softmax_rows_long_kernel<<<N, 256, 0, st>>>(S, P, Lk, Np, Np, scale * 1.4426950408889634f);
        CHECK_LAUNCH("softmax_rows");
        r = b200_gemm_bf16(P, kv + 2 * C, reinterpret_cast<__nv_bfloat16*>(out) + (long long)f * N * C, N, C, Lk8, Np, 3LL * C, C,
                           nullptr, nullptr, nullptr, 0, 0, 0, 1, stream);
        if (r) return r;
    }
    return B200_OK;
}
