// Micro-benchmark: issue rate and latency of the warp-level mma.sync.m16n8k8 TF32 on B200 (the legacy tensor path,
// SASS HMMA.1688.F32.TF32), the building block of the three-body kernels' small per-node products.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32 mma_tf32.cu && ./mma_tf32
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void k(float* out, int iters, unsigned seed) {
  float acc[CHAINS][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + (threadIdx.x + seed + i) * 1e-3f);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + (threadIdx.x + i) * 1e-3f);
  for (int c = 0; c < CHAINS; ++c)
    for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma_tf32(acc[c], a, b);
  }
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c)
    for (int i = 0; i < 4; ++i) s += acc[c][i];
  if (s == 12345.678f) out[0] = s;
}

template <int CHAINS>
void run(int warps_per_sm, int iters) {
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int threads = 128, blocks = 148 * warps_per_sm / 4;
  k<CHAINS><<<blocks, threads>>>(out, 10, 1);
  cudaEventRecord(e0);
  k<CHAINS><<<blocks, threads>>>(out, iters, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * 4 * iters * CHAINS;
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double per_sm_clk = n / (ms * 1e-3) / 148.0 / (clk_khz * 1e3);
  printf("chains %d warps/SM %2d: %.3f ms, %.2f Gmma/s, %.3f mma/clk/SM (at %d MHz nominal), %.1f TFLOP/s dense-equivalent\n",
         CHAINS, warps_per_sm, ms, n / ms * 1e-6, per_sm_clk, clk_khz / 1000, n * 2 * 16 * 8 * 8 / ms * 1e-9);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<1>(w, 20000);
    run<4>(w, 20000);
    run<8>(w, 20000);
  }
  return 0;
}
