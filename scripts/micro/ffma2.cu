// FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a: 8 independent accumulator chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long o;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o) : "l"(a), "l"(b), "l"(c));
  return o;
}
__global__ void k1(float* out, int iters, float s) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f);
  float r = 0;
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k2(float* out, int iters, float s) {
  unsigned long long a[8];
  float2 sv = make_float2(s, s), cv = make_float2(0.5f, 0.5f);
  unsigned long long sp = *reinterpret_cast<unsigned long long*>(&sv), cp = *reinterpret_cast<unsigned long long*>(&cv);
  for (int i = 0; i < 8; ++i) { float2 t = make_float2(threadIdx.x * 0.001f + i, i); a[i] = *reinterpret_cast<unsigned long long*>(&t); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], sp, cp);
  float r = 0;
  for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&a[i]); r += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); k1<<<148 * 8, 256>>>(out, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fma1 = 148.0 * 8 * 256 * 16.0 * iters;
    printf("FFMA : %.3f ms  %.1f TFMA/s\n", ms, fma1 / ms / 1e9);
    cudaEventRecord(e0); k2<<<148 * 8, 256>>>(out, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2: %.3f ms  %.1f TFMA/s\n", ms, fma1 / ms / 1e9);
  }
  return 0;
}
