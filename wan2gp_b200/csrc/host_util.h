// Host-side helpers shared by the translation units of libwan2gp_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace b200 { struct GemmParams; }

int b200_set_error(int code, const char* fmt, ...);
void b200_count_launch();
int b200_num_sms();
bool b200_first_use_on_device(std::atomic<unsigned long long>& mask);
void b200_mark_used_on_device(std::atomic<unsigned long long>& mask);
int b200_make_tmap_bf16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes);
int b200_launch_gemm_k96(int BN, const CUtensorMap& ta, const CUtensorMap& tb, const b200::GemmParams& p, cudaStream_t st);
int b200_launch_gemm(int BN, bool mn, const CUtensorMap& ta, const CUtensorMap& tb, const b200::GemmParams& p, cudaStream_t st);
int b200_pick_bn(int N, bool mn);
