// Micro-benchmark: latency and throughput of row gathers into shared memory on B200, three mechanisms:
//   mode 0: cp.async.bulk (TMA unit, one instruction per row, mbarrier completion), issued by one lane per row
//   mode 1: cp.async 16 B (LDGSTS) issued by all lanes, completion by wait_group
//   mode 2: plain LDG.128 into registers + STS
// Each warp repeatedly fetches ROWS random rows of ROWB bytes and waits for them (no compute).  Reports round latency
// (cycles) and aggregate GB/s as a function of warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy bulk_copy.cu && ./bulk_copy
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <random>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(const float* __restrict__ src, const int* __restrict__ rowid, int nrows_total, int ROWS, int ROWB,
                  int rounds, int smem_per_warp, long long* cycles, float* sink) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* my = sm + 64 + warp * smem_per_warp;
  const uint32_t bar = smem_u32(sm + warp * 8);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  __syncwarp();
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  uint32_t phase = 0;
  float acc = 0.f;
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const int base = ((gw * rounds + r) * ROWS) % (nrows_total - ROWS);
    if (MODE == 0) {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWS * ROWB) : "memory");
      __syncwarp();
      if (lane < ROWS) {
        const int row = rowid[base + lane];
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(my + lane * ROWB)),
                     "l"(reinterpret_cast<const char*>(src) + (size_t)row * ROWB), "r"(ROWB), "r"(bar)
                     : "memory");
      }
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      phase ^= 1;
    } else if (MODE == 1) {
      for (int i = 0; i < ROWS; ++i) {
        const int row = rowid[base + i];
        const char* p = reinterpret_cast<const char*>(src) + (size_t)row * ROWB;
        for (int o = lane * 16; o < ROWB; o += 512)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my + i * ROWB + o)), "l"(p + o) : "memory");
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    } else {
      for (int i = 0; i < ROWS; ++i) {
        const int row = rowid[base + i];
        const char* p = reinterpret_cast<const char*>(src) + (size_t)row * ROWB;
        for (int o = lane * 16; o < ROWB; o += 512) {
          float4 v = __ldg(reinterpret_cast<const float4*>(p + o));
          *reinterpret_cast<float4*>(my + i * ROWB + o) = v;
        }
      }
      __syncwarp();
    }
    acc += *reinterpret_cast<float*>(my + (lane * 4) % (ROWS * ROWB));
  }
  long long t1 = clock64();
  if (lane == 0) cycles[gw] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const int ROWB = 2048, NR = 400000;  // 800 MB source: rows of 2 KB (B rows 1536 + gate 512)
  float* src; int* rowid; long long* cyc; float* sink;
  cudaMalloc(&src, (size_t)NR * ROWB); cudaMemset(src, 0, (size_t)NR * ROWB);
  std::vector<int> ids(NR);
  for (int i = 0; i < NR; ++i) ids[i] = i;
  std::mt19937 rng(1); std::shuffle(ids.begin(), ids.end(), rng);
  cudaMalloc(&rowid, NR * 4); cudaMemcpy(rowid, ids.data(), NR * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&cyc, 148 * 64 * 8); cudaMalloc(&sink, 4);
  for (int mode = 0; mode < 3; ++mode)
    for (int rows : {8, 16})
      for (int wps : {4, 8, 16}) {  // warps per SM (one CTA of 4 warps x wps/4 CTAs)
        const int smem_per_warp = rows * ROWB;
        const int ctas_per_sm = wps / 4, threads = 128;
        const size_t smem = 64 + 4 * (size_t)smem_per_warp;
        if (smem * ctas_per_sm > 220 * 1024) continue;
        const int rounds = 64;
        auto fn = mode == 0 ? k<0> : mode == 1 ? k<1> : k<2>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int blocks = 148 * ctas_per_sm;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        fn<<<blocks, threads, smem>>>(src, rowid, NR, rows, ROWB, 4, smem_per_warp, cyc, sink);
        cudaEventRecord(e0);
        fn<<<blocks, threads, smem>>>(src, rowid, NR, rows, ROWB, rounds, smem_per_warp, cyc, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(blocks * 4);
        cudaMemcpy(h.data(), cyc, blocks * 4 * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto c : h) avg += c; avg /= h.size();
        const double bytes = (double)blocks * 4 * rounds * rows * ROWB;
        printf("mode %d rows/round %2d warps/SM %2d: %.3f ms, %.0f GB/s, %.0f cycles per round (err %s)\n", mode, rows, wps, ms,
               bytes / ms * 1e-6, avg / rounds, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
