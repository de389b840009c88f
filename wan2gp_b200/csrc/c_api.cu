// C ABI (include/wan2gp_b200.h) -> kernel launches.  Host-side work here is only: argument checks,
// TMA descriptor encoding (cuTensorMapEncodeTiled via the runtime's driver entry point, so the library
// has no link-time dependency on libcuda and loads on a CPU-only box), grid sizing, launch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/wan2gp_b200.h"
#include "attn2_sm100.cuh"
#include "attn3_sm100.cuh"
#include "attn5_sm100.cuh"
#include "attn6_sm100.cuh"
#include "attn_sm100.cuh"
#include "t5_ops.cuh"
#include "llm_ops.cuh"
#include "elementwise.cuh"
#include "gemm2_sm100.cuh"
#include "gemm_sm100.cuh"
#include "host_util.h"

using namespace b200;

// ------------------------------------------------------------------ error / bookkeeping
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int b200_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void b200_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" const char* b200_last_error(void) { return g_err; }
extern "C" int b200_version(void) { return 200; }
static_assert(CFG_DOTS_FLOATS == B200_CFG_DOTS_FLOATS, "header constant out of sync with the kernel");
extern "C" long long b200_launch_count(void) { return g_launches.load(); }

// Per-device caches (a process may drive several devices, one host thread each): SM count and "max dynamic shared memory
// attribute already set for kernel X on device d" bit masks.  Relaxed atomics: the cached values are idempotent.
int b200_num_sms() {
    static std::atomic<int> sms[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev & 63;
    int v = sms[slot].load(std::memory_order_relaxed);
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        sms[slot].store(v, std::memory_order_relaxed);
    }
    return v;
}
// true the first time it is called for (mask, current device): the caller then sets the per-device function attribute
bool b200_first_use_on_device(std::atomic<unsigned long long>& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return !(mask.load(std::memory_order_relaxed) & bit);
}
void b200_mark_used_on_device(std::atomic<unsigned long long>& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    mask.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
}

// ------------------------------------------------------------------ tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// bf16 tensor, dims[0] is the contiguous dimension; strides (bytes) for dims 1..rank-1; 128B swizzle.
int b200_make_tmap(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, int is_f32);
int b200_make_tmap_bf16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes) {
    return b200_make_tmap(m, ptr, rank, dims, strides_bytes, box, swizzle_bytes, 0);
}
int b200_make_tmap(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, int is_f32) {
    auto enc = get_encode();
    if (!enc) return b200_set_error(B200_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(m, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return b200_set_error(B200_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,..] box=[%u,%u,..]", (int)r,
                              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                              rank > 1 ? box[1] : 0);
    return B200_OK;
}

#define CHECK_LAUNCH(name)                                                                                   \
    do {                                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                                \
        if (e__ != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(e__)); \
        b200_count_launch();                                                                                 \
    } while (0)

// ------------------------------------------------------------------ GEMM
template <int BN, bool MN, int BKC = 64, int NBOX = 1, bool EPI_TMA = false>
static int launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st,
                            const CUtensorMap* tc = nullptr) {
    auto kern = gemm_tcgen05_kernel<BN, MN, BKC, NBOX, EPI_TMA>;
    static std::atomic<unsigned long long> attr_done{0};
    if (b200_first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN, BKC, NBOX, EPI_TMA>::kBytes);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "gemm smem attr: %s", cudaGetErrorString(e));
        b200_mark_used_on_device(attr_done);
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < b200_num_sms() ? tiles : b200_num_sms();
    kern<<<grid, 256, GemmSmem<BN, BKC, NBOX, EPI_TMA>::kBytes, st>>>(ta, tb, tc ? *tc : ta, p);
    CHECK_LAUNCH("gemm_tcgen05");
    return B200_OK;
}

// CTA-pair (cta_group::2) kernel: 256 x 256 tiles, one cluster of 2 CTAs per tile, grid = 2 * min(#tiles, #SMs / 2)
template <bool EPI_TMA>
static int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st, const CUtensorMap* tc = nullptr) {
    auto kern = gemm_pair_tcgen05_kernel<EPI_TMA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem<EPI_TMA>::kBytes);
    if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "gemm pair smem attr: %s", cudaGetErrorString(e));
    const int tiles = p.m_tiles * p.n_tiles;
    const int pairs = tiles < b200_num_sms() / 2 ? tiles : b200_num_sms() / 2;
    kern<<<2 * pairs, 256, Gemm2Smem<EPI_TMA>::kBytes, st>>>(ta, tb, tc ? *tc : ta, p);
    CHECK_LAUNCH("gemm_pair_tcgen05");
    return B200_OK;
}

// conv layers with Cin = 96: three 32-channel boxes (64B swizzle) per tap => K = 96 exactly, no zero-padded MMAs
int b200_launch_gemm_k96(int BN, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    switch (BN) {
        case 96: return launch_gemm_inst<96, false, 32, 3>(ta, tb, p, st);
        case 32: return launch_gemm_inst<32, false, 32, 3>(ta, tb, p, st);        // tap-stacked decoder head (27 -> 32 channels)
        case 16: return launch_gemm_inst<16, false, 32, 3>(ta, tb, p, st);
    }
    return b200_set_error(B200_ERR_ARG, "no K=96 GEMM instance for BN=%d", BN);
}

int b200_launch_gemm(int BN, bool mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    if (mn) {
        switch (BN) {
            case 256: return launch_gemm_inst<256, true>(ta, tb, p, st);
            case 128: return launch_gemm_inst<128, true>(ta, tb, p, st);
            case 64: return launch_gemm_inst<64, true>(ta, tb, p, st);
        }
    } else {
        switch (BN) {
            case 256: return launch_gemm_inst<256, false>(ta, tb, p, st);
            case 192: return launch_gemm_inst<192, false>(ta, tb, p, st);
            case 128: return launch_gemm_inst<128, false>(ta, tb, p, st);
            case 96: return launch_gemm_inst<96, false>(ta, tb, p, st);
            case 64: return launch_gemm_inst<64, false>(ta, tb, p, st);
            case 32: return launch_gemm_inst<32, false>(ta, tb, p, st);
            case 16: return launch_gemm_inst<16, false>(ta, tb, p, st);
        }
    }
    return b200_set_error(B200_ERR_ARG, "no GEMM instance for BN=%d mn=%d", BN, (int)mn);
}

// N tile: minimise padded columns, with a mild preference for wide tiles (fewer A re-reads, better MMA shape)
int b200_pick_bn(int N, bool mn) {
    const int cands_k[] = {256, 192, 128, 96, 64, 32, 16};
    const int cands_mn[] = {256, 128, 64};
    const int* c = mn ? cands_mn : cands_k;
    const int n = mn ? 3 : 7;
    int best = c[0];
    double best_cost = 1e30;
    for (int i = 0; i < n; ++i) {
        const double padded = (double)((N + c[i] - 1) / c[i]) * c[i];
        const double cost = padded * (1.0 + 32.0 / c[i]);
        if (cost < best_cost) { best_cost = cost; best = c[i]; }
    }
    return best;
}

extern "C" int b200_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, long long lda, long long ldb,
                              long long ldc, const float* bias, const float* gate, const void* residual_bf16, int act,
                              int out_fp32, int accumulate, int b_mn_major, void* stream) {
    if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0) return b200_set_error(B200_ERR_ARG, "gemm: null/empty argument");
    if (N % 8 || K % 8 || lda % 8 || ldb % 8 || ldc % 8) return b200_set_error(B200_ERR_ARG, "gemm: N%%8, K%%8, ld%%8 required (N=%d K=%d)", N, K);
    if (accumulate && !out_fp32) return b200_set_error(B200_ERR_ARG, "gemm: accumulate needs fp32 out");
    if (b_mn_major && N % 64) return b200_set_error(B200_ERR_ARG, "gemm: MN-major B needs N%%64==0");
    const bool mn = b_mn_major != 0;
    const int BN = b200_pick_bn(N, mn);
    // large K-major linear layers: CTA-pair kernel (B200_GEMM_PAIR=0 selects the single-CTA 128 x 256 kernel for A/B runs)
    static int use_pair = -1;
    if (use_pair < 0) { const char* ev = getenv("B200_GEMM_PAIR"); use_pair = ev ? atoi(ev) : 1; }
    if (use_pair && !mn && BN == 256 && N % 256 == 0 && M >= 512) {
        CUtensorMap ta2, tb2, tc2;
        uint64_t da[2] = {(uint64_t)K, (uint64_t)M}, sa[1] = {(uint64_t)lda * 2};
        uint64_t db[2] = {(uint64_t)K, (uint64_t)N}, sb[1] = {(uint64_t)ldb * 2};
        uint32_t box[2] = {GEMM_BK, GEMM_BM};                 // A: 64 x 128 rows per CTA; B: 64 x 128 weight rows per CTA
        int r = b200_make_tmap_bf16(&ta2, A, 2, da, sa, box, 128);
        if (r) return r;
        r = b200_make_tmap_bf16(&tb2, B, 2, db, sb, box, 128);
        if (r) return r;
        GemmParams p;
        memset(&p, 0, sizeof(p));
        p.M = M; p.N = N; p.K = K;
        p.mode = MODE_LINEAR;
        p.num_k_iters = (K + GEMM_BK - 1) / GEMM_BK;
        p.m_tiles = (M + GEMM2_BM - 1) / GEMM2_BM;
        p.n_tiles = N / GEMM2_BN;
        // N tiles walked in groups so that a slab of weights is reused from L2 while A streams (B200_GEMM_NGROUP: A/B measurements)
        static int pair_ngroup = -1;
        if (pair_ngroup < 0) { const char* ev = getenv("B200_GEMM_NGROUP"); pair_ngroup = ev && atoi(ev) > 0 ? atoi(ev) : 16; }
        p.n_group = pair_ngroup;
        p.out = out; p.out_fp32 = out_fp32; p.accumulate = accumulate; p.ldc = ldc;
        p.bias = bias; p.gate = gate; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual_bf16); p.act = act;
        static int pair_tma_reduce = -1;
        if (pair_tma_reduce < 0) { const char* ev = getenv("B200_GEMM_TMA_REDUCE"); pair_tma_reduce = ev ? atoi(ev) : 1; }
        if (pair_tma_reduce && accumulate && out_fp32 && !residual_bf16 && act == 0 && ldc % 4 == 0) {
            uint64_t dc[2] = {(uint64_t)N, (uint64_t)M}, sc[1] = {(uint64_t)ldc * 4};
            uint32_t boxc[2] = {32, GEMM_BM};
            r = b200_make_tmap(&tc2, out, 2, dc, sc, boxc, 128, 1);
            if (r) return r;
            p.tma_reduce = 1;
            return launch_gemm_pair<true>(ta2, tb2, p, (cudaStream_t)stream, &tc2);
        }
        return launch_gemm_pair<false>(ta2, tb2, p, (cudaStream_t)stream);
    }
    CUtensorMap ta, tb;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)lda * 2};
        uint32_t box[2] = {GEMM_BK, GEMM_BM};
        int r = b200_make_tmap_bf16(&ta, A, 2, dims, str, box, 128);
        if (r) return r;
    }
    if (!mn) {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)ldb * 2};
        uint32_t box[2] = {GEMM_BK, (uint32_t)BN};
        int r = b200_make_tmap_bf16(&tb, B, 2, dims, str, box, 128);
        if (r) return r;
    } else {
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
        uint64_t str[1] = {(uint64_t)ldb * 2};
        uint32_t box[2] = {64, GEMM_BK};
        int r = b200_make_tmap_bf16(&tb, B, 2, dims, str, box, 128);
        if (r) return r;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K;
    p.mode = MODE_LINEAR;
    p.num_k_iters = (K + GEMM_BK - 1) / GEMM_BK;
    p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    p.n_tiles = (N + BN - 1) / BN;
    p.n_group = 16;
    p.out = out; p.out_fp32 = out_fp32; p.accumulate = accumulate; p.ldc = ldc;
    p.bias = bias; p.gate = gate; p.residual = reinterpret_cast<const __nv_bfloat16*>(residual_bf16); p.act = act;
    // residual-stream accumulate: L2-side reduce-add through TMA (B200_GEMM_TMA_REDUCE=0 selects the LDG/STG epilogue)
    static int use_tma_reduce = -1;
    if (use_tma_reduce < 0) { const char* ev = getenv("B200_GEMM_TMA_REDUCE"); use_tma_reduce = ev ? atoi(ev) : 1; }
    if (use_tma_reduce && accumulate && out_fp32 && !mn && BN == 256 && N % 32 == 0 && !residual_bf16 && act == 0 && ldc % 4 == 0) {
        CUtensorMap tc;
        uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)ldc * 4};
        uint32_t box[2] = {32, GEMM_BM};
        int r = b200_make_tmap(&tc, out, 2, dims, str, box, 128, 1);
        if (r) return r;
        p.tma_reduce = 1;
        return launch_gemm_inst<256, false, 64, 1, true>(ta, tb, p, (cudaStream_t)stream, &tc);
    }
    return b200_launch_gemm(BN, mn, ta, tb, p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ attention
static int attention_impl(const void* q, const void* k, const void* v, void* out, int nseq, int Lq, int Lk, int H,
                          long long ldq, long long ldk, long long ldv, long long ldo, float scale, void* stream);
extern "C" int b200_attention_d128(const void* q, const void* k, const void* v, void* out, int Lq, int Lk, int H,
                                   long long ldq, long long ldk, long long ldv, long long ldo, float scale, void* stream) {
    return attention_impl(q, k, v, out, 1, Lq, Lk, H, ldq, ldk, ldv, ldo, scale, stream);
}
// nseq equally long sequences stacked along the rows: q / out [nseq * Lq, H*128], k / v [nseq * Lk, H*128]; sequence z attends only to
// its own keys.  One launch for the cond / uncond branches of a CFG pair (any2video.py:1625-1646 runs them as two forwards).
extern "C" int b200_attention_d128_batched(const void* q, const void* k, const void* v, void* out, int nseq, int Lq, int Lk, int H,
                                           long long ldq, long long ldk, long long ldv, long long ldo, float scale, void* stream) {
    if (nseq < 1 || nseq > 65535) return b200_set_error(B200_ERR_ARG, "attention_batched: nseq out of range");
    return attention_impl(q, k, v, out, nseq, Lq, Lk, H, ldq, ldk, ldv, ldo, scale, stream);
}
static int attention_impl(const void* q, const void* k, const void* v, void* out, int nseq, int Lq, int Lk, int H,
                          long long ldq, long long ldk, long long ldv, long long ldo, float scale, void* stream) {
    if (!q || !k || !v || !out || Lq <= 0 || Lk <= 0 || H <= 0) return b200_set_error(B200_ERR_ARG, "attention: null/empty argument");
    if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return b200_set_error(B200_ERR_ARG, "attention: row strides must be multiples of 8");
    CUtensorMap tq, tk, tv;
    uint32_t box[2] = {64, 128};
    {
        uint64_t dims[2] = {(uint64_t)H * 128, (uint64_t)Lq * nseq}; uint64_t str[1] = {(uint64_t)ldq * 2};
        int r = b200_make_tmap_bf16(&tq, q, 2, dims, str, box, 128); if (r) return r;
    }
    {
        uint64_t dims[2] = {(uint64_t)H * 128, (uint64_t)Lk * nseq}; uint64_t str[1] = {(uint64_t)ldk * 2};
        int r = b200_make_tmap_bf16(&tk, k, 2, dims, str, box, 128); if (r) return r;
    }
    {
        uint64_t dims[2] = {(uint64_t)H * 128, (uint64_t)Lk * nseq}; uint64_t str[1] = {(uint64_t)ldv * 2};
        int r = b200_make_tmap_bf16(&tv, v, 2, dims, str, box, 128); if (r) return r;
    }
    AttnParams p;
    p.Lq = Lq; p.Lk = Lk; p.H = H;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo;
    p.scale_log2 = scale * 1.4426950408889634f;
    // long query sequences: CTA-pair kernel (attn2_sm100.cuh): 512 query rows per cluster, each CTA stages half of every K/V tile.
    // B200_ATT_PAIR=0 selects the single-CTA kernel (A/B runs); other values: see below.
    static int use_pair = -1;
    if (use_pair < 0) { const char* ev = getenv("B200_ATT_PAIR"); use_pair = ev ? atoi(ev) : 0; }
    if (use_pair && Lq >= 1024 && nseq == 1) {
        CUtensorMap tk2;
        uint32_t boxk[2] = {64, 64};                         // this CTA's 64 keys x one 64-wide d slab
        uint64_t dims[2] = {(uint64_t)H * 128, (uint64_t)Lk}; uint64_t str[1] = {(uint64_t)ldk * 2};
        int r = b200_make_tmap_bf16(&tk2, k, 2, dims, str, boxk, 128); if (r) return r;
        dim3 grid2(2 * ((Lq + 2 * ATT_QTILES * ATT_BM - 1) / (2 * ATT_QTILES * ATT_BM)), H);
        auto launch2 = [&](auto kern) -> int {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM_BYTES);
            if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "attention pair smem attr: %s", cudaGetErrorString(e));
            kern<<<grid2, ATT_THREADS, ATT2_SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk2, tv, p);
            return B200_OK;
        };
        // variants (B200_ATT_PAIR): 1 = packed-fp32 softmax (default), 2 = scalar softmax, 4 = packed + every 4th exp2 on the FMA pipe
        int rc2 = use_pair == 4 ? launch2(attn_pair_fwd_d128_kernel<4, true>)
                : use_pair == 2 ? launch2(attn_pair_fwd_d128_kernel<0, false>) : launch2(attn_pair_fwd_d128_kernel<0, true>);
        if (rc2) return rc2;
        CHECK_LAUNCH("attn_pair_fwd_d128");
        return B200_OK;
    }
    dim3 grid((Lq + ATT_QTILES * ATT_BM - 1) / (ATT_QTILES * ATT_BM), H, nseq);
    // tuning variants (B200_ATT_VARIANT); 1 = the round-1 kernel (scalar softmax, all exp2 on MUFU, split P)
    static int variant = -1;
    if (variant < 0) {
        const char* ev = getenv("B200_ATT_VARIANT");
        // default 614 = attn6_sm100.cuh: one Q tile per CTA, three score buffers, alternating softmax warpgroups, 1/4 of the exp2 pairs on the
        // FMA pipe, one P-store wait per tile (profiles/attn_variants_r02_call11.json: 80.2 ms against 88.5 ms for 103, the two-Q-tile kernel)
        variant = ev ? atoi(ev) : 614;
    }
    auto launch = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "attention smem attr: %s", cudaGetErrorString(e));
        kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
        return B200_OK;
    };
    auto launch3 = [&](auto kern) -> int {        // 16 softmax warps (attn3_sm100.cuh)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT3_SMEM_BYTES);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "attention smem attr: %s", cudaGetErrorString(e));
        kern<<<grid, ATT3_THREADS, ATT3_SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
        return B200_OK;
    };
    auto launch_mc = [&](auto kern) -> int {      // clusters of two adjacent Q blocks sharing every K/V tile through TMA multicast
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "attention smem attr: %s", cudaGetErrorString(e));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((grid.x + 1) & ~1u, grid.y, grid.z);      // an odd block count gets one all-out-of-range Q block
        cfg.blockDim = dim3(ATT_THREADS);
        cfg.dynamicSmemBytes = ATT_SMEM_BYTES;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, p);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "attention cluster launch: %s", cudaGetErrorString(e));
        return B200_OK;
    };
    int att5_smem = ATT5_SMEM_BYTES;
    auto launch5 = [&](auto kern) -> int {        // attn5 / attn6_sm100.cuh: one Q tile per CTA, clusters of two sharing K/V
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, att5_smem);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "attention smem attr: %s", cudaGetErrorString(e));
        const unsigned q_tiles = (unsigned)((Lq + ATT_BM - 1) / ATT_BM);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((q_tiles + 1) & ~1u, grid.y, grid.z);     // an odd tile count gets one all-out-of-range Q tile
        cfg.blockDim = dim3(ATT5_THREADS);
        cfg.dynamicSmemBytes = att5_smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, p);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "attention cluster launch: %s", cudaGetErrorString(e));
        return B200_OK;
    };
    int rc;
    switch (variant) {
        // one Q tile per CTA with double-buffered scores (S_{j+1} runs under the softmax of tile j); 50x: every x-th exp2 pair on the FMA pipe
        // three score buffers, two softmax warpgroups alternating over the K/V tiles, fixed reference maximum (attn6_sm100.cuh)
        // 60x: P published per 64-key half; 61x: both halves after one wait
        case 600: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<0, true>); break;
        case 603: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<3, true>); break;
        case 604: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<4, true>); break;
        case 613: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<3, false>); break;
        case 612: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<2, false>); break;
        case 615: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<5, false>); break;
        case 616: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<6, false>); break;
        case 610: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<0, false>); break;
        case 500: rc = launch5(attn_s2_fwd_d128_kernel<0>); break;
        case 503: rc = launch5(attn_s2_fwd_d128_kernel<3>); break;
        case 504: rc = launch5(attn_s2_fwd_d128_kernel<4>); break;
        case 502: rc = launch5(attn_s2_fwd_d128_kernel<2>); break;
        // K/V tiles multicast to a cluster of two CTAs (each loads half): 300 = packed softmax, 303 = + every 3rd exp2 pair on the FMA pipe
        case 300: rc = launch_mc(attn_fwd_d128_kernel<0, true, true, true>); break;
        case 303: rc = launch_mc(attn_fwd_d128_kernel<3, true, true, true>); break;
        // two softmax warpgroups per Q tile (one per 64-key half), packed-fp32 softmax; 203: every 3rd exp2 pair on the FMA pipe
        case 200: rc = launch3(attn_fwd_d128_w16_kernel<0>); break;
        case 203: rc = launch3(attn_fwd_d128_w16_kernel<3>); break;
        case 204: rc = launch3(attn_fwd_d128_w16_kernel<4>); break;
        // measured at L=75600, H=40 (profiles/attn_variants_r01.txt): 0: 100.5 ms, 1: 97.3 ms, 41: 98.6 ms, 21: 109.0 ms --
        // under the 1 kW power cap extra FMA-pipe work for exp2 costs more clock than the MUFU relief buys
        case 0: rc = launch(attn_fwd_d128_kernel<0, false>); break;
        case 41: rc = launch(attn_fwd_d128_kernel<4, true>); break;
        case 21: rc = launch(attn_fwd_d128_kernel<2, true>); break;
        // packed-fp32 softmax (FFMA2 / FADD2 / FMNMX3); 10x: every x-th PAIR of exponentials on the FMA pipe
        case 100: rc = launch(attn_fwd_d128_kernel<0, true, true>); break;
        case 104: rc = launch(attn_fwd_d128_kernel<4, true, true>); break;
        case 103: rc = launch(attn_fwd_d128_kernel<3, true, true>); break;
        case 102: rc = launch(attn_fwd_d128_kernel<2, true, true>); break;
        case 1: rc = launch(attn_fwd_d128_kernel<0, true>); break;
        // limiter ablations of the default kernel: timing only, the output is wrong by construction (attn_sm100.cuh "ABL")
        case 901: rc = launch(attn_fwd_d128_kernel<3, true, true, false, 1>); break;
        case 902: rc = launch(attn_fwd_d128_kernel<3, true, true, false, 2>); break;
        case 903: rc = launch(attn_fwd_d128_kernel<3, true, true, false, 3>); break;
        case 904: rc = launch(attn_fwd_d128_kernel<3, true, true, false, 4>); break;
        case 905: rc = launch(attn_fwd_d128_kernel<3, true, true, false, 5>); break;
        case 614:
        default: att5_smem = ATT6_SMEM_BYTES; rc = launch5(attn_s3_fwd_d128_kernel<4, false>); break;
    }
    if (rc) return rc;
    CHECK_LAUNCH("attn_fwd_d128");
    return B200_OK;
}

// ------------------------------------------------------------------ row / elementwise kernels
extern "C" int b200_ln_modulate(const float* x, const float* shift, const float* scale, int affine, int pre_round, void* y,
                                int L, int D, float eps, void* stream) {
    if (!x || !shift || !scale || !y || L <= 0) return b200_set_error(B200_ERR_ARG, "ln_modulate: null/empty argument");
    if (D % 4 || D > 256 * 4 * LN_MAXV) return b200_set_error(B200_ERR_ARG, "ln_modulate: D=%d unsupported", D);
    ln_modulate_kernel<<<L, 256, 0, (cudaStream_t)stream>>>(x, shift, scale, affine, pre_round, reinterpret_cast<__nv_bfloat16*>(y), D, eps);
    CHECK_LAUNCH("ln_modulate");
    return B200_OK;
}

// B200_RMSROPE_PIPE (default 1): full-width rows through the cp.async-pipelined kernel; 0 = one CTA per row (A/B runs).  Per-head norms
// (Hunyuan) and short inputs keep the one-CTA-per-row kernel.
static bool rmsrope_pipe_wanted(int L, int per_head, const void* a, const void* b) {
    static int use = -1;
    if (use < 0) { const char* ev = getenv("B200_RMSROPE_PIPE"); use = ev ? atoi(ev) : 1; }
    return use && !per_head && L >= 2048 && !(((uintptr_t)a | (uintptr_t)b) & 15);
}
static int launch_rmsrope_pipe(__nv_bfloat16* a, __nv_bfloat16* b, long long ld, const float* wa, const float* wb, int L, int D, float eps,
                               const float* cos_t, const float* sin_t, int nseg, void* stream) {
    const long long n_items = (long long)L * nseg;
    const size_t smem = (size_t)RP_STAGES * (D >> 3) * sizeof(uint4);           // 3 rows: 30 KB at D = 5120 (below the 48 KB default limit)
    if (smem > 40 * 1024) return -1;     // D > 6826: the ring + the reduction scratch would pass the 48 KB default limit -> one-CTA-per-row kernel
    rmsnorm_rope_pipe_kernel<<<(unsigned)((n_items + RP_ITEMS - 1) / RP_ITEMS), 256, smem, (cudaStream_t)stream>>>(a, b, ld, wa, wb, D, eps, cos_t, sin_t,
                                                                                                                 n_items, nseg);
    return 0;
}

extern "C" int b200_rmsnorm_rope(void* x, long long ld, const float* w, int L, int D, float eps, const float* cos_t,
                                 const float* sin_t, int per_head, void* stream) {
    if (!x || !w || L <= 0) return b200_set_error(B200_ERR_ARG, "rmsnorm_rope: null/empty argument");
    if (D % 128 || D > 256 * 8 * RN_MAXV || ld % 8) return b200_set_error(B200_ERR_ARG, "rmsnorm_rope: D=%d ld=%lld unsupported", D, ld);
    if ((cos_t == nullptr) != (sin_t == nullptr)) return b200_set_error(B200_ERR_ARG, "rmsnorm_rope: cos/sin must both be given");
    auto xb = reinterpret_cast<__nv_bfloat16*>(x);
    if (rmsrope_pipe_wanted(L, per_head, x, x) && launch_rmsrope_pipe(xb, xb, ld, w, w, L, D, eps, cos_t, sin_t, 1, stream) == 0) {
        CHECK_LAUNCH("rmsnorm_rope_pipe");
        return B200_OK;
    }
    if (per_head) rmsnorm_rope_kernel<true><<<L, 256, 0, (cudaStream_t)stream>>>(xb, xb, ld, w, w, D, eps, cos_t, sin_t);
    else rmsnorm_rope_kernel<false><<<L, 256, 0, (cudaStream_t)stream>>>(xb, xb, ld, w, w, D, eps, cos_t, sin_t);
    CHECK_LAUNCH("rmsnorm_rope");
    return B200_OK;
}

extern "C" int b200_qk_rmsnorm_rope(void* q, void* k, long long ld, const float* wq, const float* wk, int L, int D, float eps,
                                    const float* cos_t, const float* sin_t, int per_head, void* stream) {
    if (!q || !k || !wq || !wk || L <= 0) return b200_set_error(B200_ERR_ARG, "qk_rmsnorm_rope: null/empty argument");
    if (D % 128 || D > 256 * 8 * RN_MAXV || ld % 8) return b200_set_error(B200_ERR_ARG, "qk_rmsnorm_rope: D=%d ld=%lld unsupported", D, ld);
    if ((cos_t == nullptr) != (sin_t == nullptr)) return b200_set_error(B200_ERR_ARG, "qk_rmsnorm_rope: cos/sin must both be given");
    auto qb = reinterpret_cast<__nv_bfloat16*>(q), kb = reinterpret_cast<__nv_bfloat16*>(k);
    if (rmsrope_pipe_wanted(L, per_head, q, k) && launch_rmsrope_pipe(qb, kb, ld, wq, wk, L, D, eps, cos_t, sin_t, 2, stream) == 0) {
        CHECK_LAUNCH("qk_rmsnorm_rope_pipe");
        return B200_OK;
    }
    const dim3 grid(L, 2);
    if (per_head) rmsnorm_rope_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(qb, kb, ld, wq, wk, D, eps, cos_t, sin_t);
    else rmsnorm_rope_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(qb, kb, ld, wq, wk, D, eps, cos_t, sin_t);
    CHECK_LAUNCH("qk_rmsnorm_rope");
    return B200_OK;
}

extern "C" int b200_cast_f32_bf16(const float* x, void* y, long long n, void* stream) {
    if (!x || !y || n <= 0 || n % 4) return b200_set_error(B200_ERR_ARG, "cast: bad argument");
    const long long n4 = n / 4;
    cast_f32_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n4);
    CHECK_LAUNCH("cast_f32_bf16");
    return B200_OK;
}

extern "C" int b200_patch_embed(const float* x0, int C0, const float* x1, int C1, const float* w, const float* bias,
                                float* out, int T, int H, int W, int D, int patch, void* stream) {
    if (!x0 || !w || !bias || !out || (C1 > 0 && !x1)) return b200_set_error(B200_ERR_ARG, "patch_embed: null argument");
    if ((patch != 1 && patch != 2) || H % patch || W % patch) return b200_set_error(B200_ERR_ARG, "patch_embed: patch must be 1 or 2 and divide H, W");
    const int L = T * (H / patch) * (W / patch);
    dim3 grid((L + PE_TOK - 1) / PE_TOK, (D + PE_CH - 1) / PE_CH);
    patch_embed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x0, C0, x1, C1, w, bias, out, T, H, W, D, patch);
    CHECK_LAUNCH("patch_embed");
    return B200_OK;
}

extern "C" int b200_unpatchify(const float* y, float* out, int C, int T, int H, int W, int patch, int c_major, void* stream) {
    if (!y || !out || (patch != 1 && patch != 2)) return b200_set_error(B200_ERR_ARG, "unpatchify: bad argument");
    const long long n = (long long)C * T * H * W;
    unpatchify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, out, C, T, H, W, patch, c_major);
    CHECK_LAUNCH("unpatchify");
    return B200_OK;
}

extern "C" int b200_gemv_f32(const float* x, const float* w, const float* b, float* out, int N, int K, int silu_in,
                             int silu_out, void* stream) {
    if (!x || !w || !b || !out || K % 4) return b200_set_error(B200_ERR_ARG, "gemv: bad argument");
    gemv_f32_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, w, b, out, N, K, silu_in, silu_out);
    CHECK_LAUNCH("gemv_f32");
    return B200_OK;
}

extern "C" int b200_sinusoid(float t, float* out, int dim, void* stream) {
    if (!out || dim % 2) return b200_set_error(B200_ERR_ARG, "sinusoid: bad argument");
    sinusoid_kernel<<<(dim / 2 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, nullptr, out, dim);
    CHECK_LAUNCH("sinusoid");
    return B200_OK;
}
extern "C" int b200_sinusoid_dev(const float* t_dev, float* out, int dim, void* stream) {
    if (!t_dev || !out || dim % 2) return b200_set_error(B200_ERR_ARG, "sinusoid_dev: bad argument");
    sinusoid_kernel<<<(dim / 2 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(0.f, t_dev, out, dim);
    CHECK_LAUNCH("sinusoid_dev");
    return B200_OK;
}

extern "C" int b200_add_vec(const float* a, const float* b, float* out, int n, int bmod, void* stream) {
    if (!a || !b || !out || n <= 0 || bmod <= 0) return b200_set_error(B200_ERR_ARG, "add_vec: bad argument");
    add_vec_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, b, out, n, bmod);
    CHECK_LAUNCH("add_vec");
    return B200_OK;
}

extern "C" int b200_col_mean_f32(const float* x, float* out, int rows, int cols, void* stream) {
    if (!x || !out || rows <= 0 || cols <= 0) return b200_set_error(B200_ERR_ARG, "col_mean: bad argument");
    col_mean_kernel<<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, out, rows, cols);
    CHECK_LAUNCH("col_mean");
    return B200_OK;
}

static int cfg_unipc_impl(float* lat, const float* cond, const float* uncond, float guide, float* x_last, const float* m0, float* m1,
                          const float* coef_host8, int use_corrector, const float* params_dev, float* star_dots, long long n, void* stream) {
    if (!lat || !cond || !x_last || !m0 || !m1 || (!coef_host8 && !params_dev) || n <= 0 || n % 4 || (star_dots && !uncond))
        return b200_set_error(B200_ERR_ARG, "cfg_unipc_step: bad argument");
    const long long n4 = n / 4;
    if (star_dots) {
        // star_dots = float[B200_CFG_DOTS_FLOATS] scratch (include/wan2gp_b200.h): dots, per-CTA partials, ticket
        cudaMemsetAsync(star_dots + 2 + 2 * CFG_DOTS_MAX_BLOCKS, 0, 2 * sizeof(float), (cudaStream_t)stream);
        const unsigned blocks = (unsigned)((n4 + 255) / 256 < CFG_DOTS_MAX_BLOCKS ? (n4 + 255) / 256 : CFG_DOTS_MAX_BLOCKS);
        cfg_dots_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cond, uncond, star_dots, n4);
        CHECK_LAUNCH("cfg_dots");
    }
    UniPCCoef k{};
    if (coef_host8) k = UniPCCoef{coef_host8[0], coef_host8[1], coef_host8[2], coef_host8[3], coef_host8[4], coef_host8[5], coef_host8[6], coef_host8[7]};
    cfg_unipc_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lat, cond, uncond, guide, x_last, m0, m1, k, use_corrector,
                                                                                 params_dev, star_dots, n4);
    CHECK_LAUNCH("cfg_unipc_step");
    return B200_OK;
}
extern "C" int b200_cfg_unipc_step(float* lat, const float* cond, const float* uncond, float guide, float* x_last, const float* m0,
                                   float* m1, const float* coef_host8, int use_corrector, float* star_dots, long long n, void* stream) {
    if (!coef_host8) return b200_set_error(B200_ERR_ARG, "cfg_unipc_step: null coefficients");
    return cfg_unipc_impl(lat, cond, uncond, guide, x_last, m0, m1, coef_host8, use_corrector, nullptr, star_dots, n, stream);
}
extern "C" int b200_cfg_unipc_step_dev(float* lat, const float* cond, const float* uncond, float* x_last, const float* m0, float* m1,
                                       const float* params_dev10, float* star_dots, long long n, void* stream) {
    if (!params_dev10) return b200_set_error(B200_ERR_ARG, "cfg_unipc_step_dev: null parameter buffer");
    return cfg_unipc_impl(lat, cond, uncond, 0.f, x_last, m0, m1, nullptr, 0, params_dev10, star_dots, n, stream);
}

static int cfg_euler_impl(float* lat, const float* cond, const float* uncond, float guide, float dt, const float* gdt_dev,
                          float* pred_out, float* star_dots, long long n, void* stream);
extern "C" int b200_cfg_euler_step(float* lat, const float* cond, const float* uncond, float guide, float dt,
                                   float* pred_out, float* star_dots, long long n, void* stream) {
    return cfg_euler_impl(lat, cond, uncond, guide, dt, nullptr, pred_out, star_dots, n, stream);
}
extern "C" int b200_cfg_euler_step_dev(float* lat, const float* cond, const float* uncond, const float* guide_dt_dev,
                                       float* pred_out, float* star_dots, long long n, void* stream) {
    if (!guide_dt_dev) return b200_set_error(B200_ERR_ARG, "cfg_euler_step_dev: null parameter buffer");
    return cfg_euler_impl(lat, cond, uncond, 0.f, 0.f, guide_dt_dev, pred_out, star_dots, n, stream);
}
static int cfg_euler_impl(float* lat, const float* cond, const float* uncond, float guide, float dt, const float* gdt_dev,
                          float* pred_out, float* star_dots, long long n, void* stream) {
    if (!lat || !cond || n <= 0 || n % 4 || (star_dots && !uncond)) return b200_set_error(B200_ERR_ARG, "cfg_euler_step: bad argument");
    const long long n4 = n / 4;
    if (star_dots) {
        // star_dots = float[B200_CFG_DOTS_FLOATS] scratch (include/wan2gp_b200.h): dots, per-CTA partials, ticket
        cudaMemsetAsync(star_dots + 2 + 2 * CFG_DOTS_MAX_BLOCKS, 0, 2 * sizeof(float), (cudaStream_t)stream);
        const unsigned blocks = (unsigned)((n4 + 255) / 256 < CFG_DOTS_MAX_BLOCKS ? (n4 + 255) / 256 : CFG_DOTS_MAX_BLOCKS);
        cfg_dots_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cond, uncond, star_dots, n4);
        CHECK_LAUNCH("cfg_dots");
    }
    cfg_euler_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lat, cond, uncond, guide, dt, gdt_dev, pred_out, star_dots, n4);
    CHECK_LAUNCH("cfg_euler_step");
    return B200_OK;
}

// ------------------------------------------------------------------ umT5 text encoder (t5_ops.cuh)
extern "C" int b200_embed_rows(const long long* ids, const void* table, int table_is_bf16, float* out, int L, int dim, void* stream) {
    if (!ids || !table || !out || L <= 0 || dim <= 0 || dim % 2) return b200_set_error(B200_ERR_ARG, "embed_rows: bad argument");
    if (table_is_bf16) embed_rows_kernel<true><<<L, 256, 0, (cudaStream_t)stream>>>(ids, table, out, dim);
    else embed_rows_kernel<false><<<L, 256, 0, (cudaStream_t)stream>>>(ids, table, out, dim);
    CHECK_LAUNCH("embed_rows");
    return B200_OK;
}

extern "C" int b200_t5_rmsnorm(const float* x, const float* w, void* out, int out_fp32, int L, int dim, float eps, void* stream) {
    if (!x || !w || !out || L <= 0 || dim <= 0) return b200_set_error(B200_ERR_ARG, "t5_rmsnorm: bad argument");
    if (out_fp32) t5_rmsnorm_kernel<true><<<L, 256, 0, (cudaStream_t)stream>>>(x, w, out, dim, eps);
    else t5_rmsnorm_kernel<false><<<L, 256, 0, (cudaStream_t)stream>>>(x, w, out, dim, eps);
    CHECK_LAUNCH("t5_rmsnorm");
    return B200_OK;
}

extern "C" int b200_mul_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
    if (!a || !b || !out || n <= 0 || n % 2) return b200_set_error(B200_ERR_ARG, "mul_bf16: bad argument");
    const long long n2 = n / 2;
    mul_bf16_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat162*>(a), reinterpret_cast<const __nv_bfloat162*>(b), reinterpret_cast<__nv_bfloat162*>(out), n2);
    CHECK_LAUNCH("mul_bf16");
    return B200_OK;
}

extern "C" int b200_t5_attention(const void* q, const void* k, const void* v, long long ld, const float* bias_rel, void* out, long long ldo,
                                 int L, int heads, int n_valid, void* stream) {
    if (!q || !k || !v || !bias_rel || !out || L <= 0 || heads <= 0) return b200_set_error(B200_ERR_ARG, "t5_attention: null/empty argument");
    if (L > 512) return b200_set_error(B200_ERR_ARG, "t5_attention: at most 512 tokens (the reference's text_len)");
    if (ld % 8 || ldo % 2 || n_valid < 1 || n_valid > L) return b200_set_error(B200_ERR_ARG, "t5_attention: strides / n_valid");
    const int Lp = (L + 31) / 32 * 32;
    const size_t smem = (size_t)Lp * T5_ATT_D * 2 * 2 + (size_t)((2 * L - 1 + 3) & ~3) * 4 + (size_t)8 * Lp * 4;
    static std::atomic<unsigned long long> attr_done{0};
    if (b200_first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(t5_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return b200_set_error(B200_ERR_CUDA, "t5_attention smem attr: %s", cudaGetErrorString(e));
        b200_mark_used_on_device(attr_done);
    }
    dim3 grid((L + T5_ATT_ROWS - 1) / T5_ATT_ROWS, heads);
    t5_attention_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
        reinterpret_cast<const __nv_bfloat16*>(v), ld, bias_rel, reinterpret_cast<__nv_bfloat16*>(out), ldo, L, Lp, n_valid);
    CHECK_LAUNCH("t5_attention");
    return B200_OK;
}

// ------------------------------------------------------------------ decoder-only LLM text towers (llm_ops.cuh)
extern "C" int b200_rope_half(void* x, long long ld, const float* cos_t, const float* sin_t, int L, int nheads, void* stream) {
    if (!x || !cos_t || !sin_t || L <= 0 || nheads <= 0) return b200_set_error(B200_ERR_ARG, "rope_half: null/empty argument");
    if (ld < (long long)nheads * LLM_HD) return b200_set_error(B200_ERR_ARG, "rope_half: row stride %lld < %d heads x 128", ld, nheads);
    rope_half_kernel<<<L, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(x), ld, cos_t, sin_t, nheads);
    CHECK_LAUNCH("rope_half");
    return B200_OK;
}

extern "C" int b200_causal_gqa_attention(const void* q, const void* k, const void* v, long long ldq, long long ldkv, void* out, long long ldo,
                                         int L, int q_heads, int kv_heads, float scale, void* stream) {
    if (!q || !k || !v || !out || L <= 0 || q_heads <= 0 || kv_heads <= 0) return b200_set_error(B200_ERR_ARG, "causal_gqa_attention: null/empty argument");
    if (q_heads % kv_heads) return b200_set_error(B200_ERR_ARG, "causal_gqa_attention: %d q heads not a multiple of %d kv heads", q_heads, kv_heads);
    if (ldq % 4 || ldkv % 8 || ldo % 4) return b200_set_error(B200_ERR_ARG, "causal_gqa_attention: strides must keep rows 8 / 16 / 8-byte aligned");
    if (((uintptr_t)q & 7) || ((uintptr_t)k & 15) || ((uintptr_t)v & 7) || ((uintptr_t)out & 7))
        return b200_set_error(B200_ERR_ARG, "causal_gqa_attention: operand alignment (q, v, out 8 bytes; k 16 bytes)");
    dim3 grid((L + 7) / 8, q_heads);
    causal_gqa_attention_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
        reinterpret_cast<const __nv_bfloat16*>(v), ldq, ldkv, reinterpret_cast<__nv_bfloat16*>(out), ldo, L, q_heads / kv_heads,
        scale * 1.4426950408889634f);
    CHECK_LAUNCH("causal_gqa_attention");
    return B200_OK;
}
