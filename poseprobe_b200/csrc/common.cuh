// Shared helpers for the poseprobe_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/poseprobe_b200.h"

namespace pp {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PP_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::pp::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

#define PP_CHECK_LAUNCH(name)                                                     \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      ::pp::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return PP_ERR_CUDA;                                                         \
    }                                                                             \
    ::pp::count_launch();                                                         \
  } while (0)

#define PP_CHECK_CUDA(expr)                                                       \
  do {                                                                            \
    cudaError_t e__ = (expr);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::pp::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));           \
      return PP_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

inline cudaStream_t as_stream(pp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// Number of SMs of the current device (148 on B200); cached.
int sm_count();

// ---- rounding-pinned arithmetic --------------------------------------------------------------
// The reference is compiled with nvcc's default -fmad=true, so `a + b*c` patterns in its kernels
// are single-rounded FMAs.  Integer outputs (N_steps, mask_outbbox, ...) depend on those roundings,
// so the product spells them out instead of trusting the contraction heuristics.
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// 128-bit streaming loads / stores (read-once data: keep it out of L1).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

}  // namespace pp
