// Shared helpers for the lcao_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "lcao_b200.h"

void lcao_set_error(const char* fmt, ...);

#define LCAO_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      lcao_set_error(__VA_ARGS__);         \
      return LCAO_E_ARG;                   \
    }                                      \
  } while (0)

#define LCAO_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t err__ = (call);                                                      \
    if (err__ != cudaSuccess) {                                                      \
      lcao_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
      return LCAO_E_CUDA;                                                            \
    }                                                                                \
  } while (0)

extern unsigned long long g_lcao_launches;  // kernels launched by this library (reported by lcao_launch_count)
#define LCAO_LAUNCH_CHECK()        \
  do {                             \
    ++g_lcao_launches;             \
    LCAO_CUDA(cudaGetLastError()); \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 8 lanes of an aligned lane-octet
__device__ __forceinline__ float octet_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf(float x) { return x * sigmoidf_acc(x); }
__device__ __forceinline__ float silu_gradf(float x) {
  float s = sigmoidf_acc(x);
  return s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// real spherical harmonics Y_l^0(c) and derivatives (shbf.py:41-63)
#define LCAO_Y0 0.28209479177387814f
#define LCAO_Y1 0.4886025119029199f
#define LCAO_Y2A 0.9461746957575601f
#define LCAO_Y2B 0.31539156525252005f
#define LCAO_Y3 0.3731763325901154f
