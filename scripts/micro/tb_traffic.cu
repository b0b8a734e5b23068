// Micro-benchmark: the memory-access floor of the three-body backward.  Same traffic as lcao_threebody_bwd on a QM9-shape
// batch (826 fully linked 18-atom molecules: E = 252 756 edges sorted by source, 17 in / out edges per node), no math:
// warp per in-edge (k->s) in in-CSR order reads B[e] (3 x 512 B) + gate[k] (512 B) + the 17 d_tbw rows of the out-edges
// of s (512 B each, mostly L1/L2 hits) and writes dB[e] (3 x 512 B) + q[e] (512 B).  MODE 1: the same rows visited in
// edge order (streaming) for comparison.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tb_traffic tb_traffic.cu && ./tb_traffic
#include <cstdio>
#include <cuda_runtime.h>

constexpr int A = 18, D = 17, C = 128, NL = 3;

template <int MODE, int GT>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ B, const float4* __restrict__ gate, const float4* __restrict__ Gt,
                                        float4* __restrict__ dB, float4* __restrict__ q, int E) {
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < E; w += nw) {
    int e, s, kn;
    if (MODE == 0) {  // in-CSR order: task w = (node s, i-th in-edge)
      s = w / D;
      const int i = w % D, mol = s / A, sl = s % A;
      const int kl = i < sl ? i : i + 1;            // source atom of the in-edge
      const int idx = sl < kl ? sl : sl - 1;        // position of s among the out-edges of k
      e = mol * A * D + kl * D + idx;
      kn = mol * A + kl;
    } else {
      e = w; s = w / D; kn = s;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g = gate[(size_t)kn * 32 + lane];
    float4 b[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) b[l] = B[((size_t)e * NL + l) * 32 + lane];
    if (GT) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const float4 x = Gt[((size_t)s * D + j) * 32 + lane];
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
    }
#pragma unroll
    for (int l = 0; l < NL; ++l)
      dB[((size_t)e * NL + l) * 32 + lane] = make_float4(b[l].x * g.x + acc.x, b[l].y * g.y + acc.y, b[l].z * g.z + acc.z, b[l].w * g.w + acc.w);
    q[(size_t)e * 32 + lane] = acc;
  }
}

int main() {
  const int M = 826, N = M * A, E = N * D;
  float4 *B, *gate, *Gt, *dB, *q, *flush;
  cudaMalloc(&B, (size_t)E * NL * 512); cudaMalloc(&dB, (size_t)E * NL * 512);
  cudaMalloc(&Gt, (size_t)E * 512); cudaMalloc(&q, (size_t)E * 512); cudaMalloc(&gate, (size_t)N * 512);
  cudaMalloc(&flush, 512u << 20);
  cudaMemset(B, 0, (size_t)E * NL * 512); cudaMemset(Gt, 0, (size_t)E * 512); cudaMemset(gate, 0, (size_t)N * 512);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)E * (NL * 512.0 * 2 + 512.0 * 2) + N * 512.0;
  for (int mode = 0; mode < 2; ++mode)
    for (int gt = 0; gt < 2; ++gt)
      for (int per_sm = 2; per_sm <= 8; per_sm *= 2) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          cudaMemset(flush, rep, 512u << 20);
          cudaEventRecord(e0);
          const int grid = 148 * per_sm;
          if (mode == 0 && gt) k<0, 1><<<grid, 256>>>(B, gate, Gt, dB, q, E);
          else if (mode == 0) k<0, 0><<<grid, 256>>>(B, gate, Gt, dB, q, E);
          else if (gt) k<1, 1><<<grid, 256>>>(B, gate, Gt, dB, q, E);
          else k<1, 0><<<grid, 256>>>(B, gate, Gt, dB, q, E);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        printf("%s order, d_tbw reads %d, %d CTAs/SM x 8 warps: %.3f ms, %.0f GB/s (B + dB + q%s)\n", mode ? "edge  " : "in-CSR", gt, per_sm, best,
               (bytes - (gt ? 0 : E * 512.0)) / best / 1e6, gt ? " + d_tbw" : "");
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
