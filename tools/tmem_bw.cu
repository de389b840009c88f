// TMEM read / write port micro-benchmark (sm_100a): how many bytes per clock can the softmax warps of the attention kernel pull out of TMEM?
// The attention K/V step moves 2 x 64 KB of fp32 scores TMEM -> registers per 2048 tensor-pipe clocks = 64 B/clk/SM -- if the read port
// tops out there, the S read alone is co-critical with the MMAs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I wan2gp_b200/csrc tools/tmem_bw.cu -o tools/tmem_bw && tools/tmem_bw
//
// Each active warp reads (or writes) the 128 columns x 32 lanes of its own lane quarter ITERS times with tcgen05.ld.32x32b.x32 (the shape
// the kernel uses) or .x16, one tcgen05.wait per 128 columns.  Warps 4..7 alias the lane quarters of warps 0..3 (two warps per SM
// sub-partition, as the two softmax warpgroups do).  Prints clocks, bytes/clk/SM for 1 SM and for all 148.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "sm100.cuh"

using namespace b200;

template <int MODE>       // 0: ld x32, 1: ld x16, 2: st x32, 3: ld x32 with the wait after every instruction, 4: ld 16x256b.x8, 5: st 16x128b.x8, 6: st 32x32b.x16
__global__ void __launch_bounds__(256, 1) tmem_bw_kernel(int iters, int nwarps, long long* clocks, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    uint32_t acc = 0;
    uint32_t v[128];
    #pragma unroll
    for (int i = 0; i < 128; ++i) v[i] = threadIdx.x + i;
    if (warp < nwarps) {
        // initialise the columns so that reads return defined data
        #pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st_32x32b_x32(base + c * 32, v + c * 32);
        tmem_st_wait();
    }
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
            if constexpr (MODE == 0) {
                #pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(base + c * 32, v + c * 32);
                tmem_ld_wait();
            } else if constexpr (MODE == 1) {
                #pragma unroll
                for (int c = 0; c < 8; ++c) tmem_ld_32x32b_x16(base + c * 16, v + c * 16);
                tmem_ld_wait();
            } else if constexpr (MODE == 2) {
                #pragma unroll
                for (int c = 0; c < 4; ++c) tmem_st_32x32b_x32(base + c * 32, v + c * 32);
                tmem_st_wait();
            } else if constexpr (MODE == 3) {
                #pragma unroll
                for (int c = 0; c < 4; ++c) { tmem_ld_32x32b_x32(base + c * 32, v + c * 32); tmem_ld_wait(); }
            } else if constexpr (MODE == 4) {       // the same 32 lanes x 128 columns as four 16-lane x 64-column fragment loads
                #pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_16x256b_x8(base + ((uint32_t)((c >> 1) * 16) << 16) + (c & 1) * 64, v + c * 32);
                tmem_ld_wait();
            } else if constexpr (MODE == 5) {       // 32 lanes x 128 columns as eight 16-lane x 32-column stores
                #pragma unroll
                for (int c = 0; c < 8; ++c) tmem_st_16x128b_x8(base + ((uint32_t)((c >> 2) * 16) << 16) + (c & 3) * 32, v + c * 16);
                tmem_st_wait();
            } else {
                #pragma unroll
                for (int c = 0; c < 8; ++c) tmem_st_32x32b_x16(base + c * 16, v + c * 16);
                tmem_st_wait();
            }
            #pragma unroll
            for (int i = 0; i < 128; i += 32) acc ^= v[i];
            if constexpr (MODE == 2 || MODE >= 5) v[0] += acc;        // constant index: the array stays in registers
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) clocks[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE>
static void run(const char* name, int grid, int nwarps, int iters, long long* d_clk, uint32_t* d_sink) {
    cudaMemset(d_clk, 0, 148 * 8 * sizeof(long long));
    tmem_bw_kernel<MODE><<<grid, 256>>>(iters, nwarps, d_clk, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    static long long h[148 * 8];
    cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int b = 0; b < grid; ++b)
        for (int w = 0; w < nwarps; ++w) mx = h[b * 8 + w] > mx ? h[b * 8 + w] : mx;
    const double bytes = (double)nwarps * iters * 128 * 32 * 4;
    printf("{\"op\": \"%s\", \"sms\": %d, \"warps\": %d, \"iters\": %d, \"clocks\": %lld, \"bytes_per_clk_per_sm\": %.2f}\n", name, grid, nwarps, iters, mx,
           bytes / (double)mx);
}

int main() {
    long long* d_clk; uint32_t* d_sink;
    cudaMalloc(&d_clk, 148 * 8 * sizeof(long long));
    cudaMalloc(&d_sink, 4);
    const int iters = 4096;
    for (int grid : {148}) {
        for (int nw : {1, 4, 8}) {
            run<0>("tcgen05.ld.32x32b.x32 (4 per wait)", grid, nw, iters, d_clk, d_sink);
            run<1>("tcgen05.ld.32x32b.x16 (8 per wait)", grid, nw, iters, d_clk, d_sink);
            run<3>("tcgen05.ld.32x32b.x32 (wait each)", grid, nw, iters, d_clk, d_sink);
            run<2>("tcgen05.st.32x32b.x32 (4 per wait)", grid, nw, iters, d_clk, d_sink);
            run<6>("tcgen05.st.32x32b.x16 (8 per wait)", grid, nw, iters, d_clk, d_sink);
            run<4>("tcgen05.ld.16x256b.x8 (4 per wait)", grid, nw, iters, d_clk, d_sink);
            run<5>("tcgen05.st.16x128b.x8 (8 per wait)", grid, nw, iters, d_clk, d_sink);
        }
    }
    return 0;
}
