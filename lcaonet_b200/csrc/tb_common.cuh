// Helpers shared by the three-body kernels (threebody.cu: FP32-pipe formulation; threebody_mma.cu: warp-level
// 3xTF32 tensor-core formulation): real spherical harmonics Y_l^0 (shbf.py:41-63), the Gram-matrix forms of the
// triplet norm, float4 arithmetic.
#pragma once
#include "common.cuh"

namespace {

constexpr float kEps = 1e-12f; // F.normalize eps (lcaonet.py:184)
constexpr int kJS = 128;       // backward (forces): out-edges whose d_unit partials live in shared memory

template <int NL>
__device__ __forceinline__ void sph_harm(float c, float (&Y)[4]) {
  Y[0] = LCAO_Y0;
  Y[1] = (NL > 1) ? LCAO_Y1 * c : 0.f;
  Y[2] = (NL > 2) ? fmaf(LCAO_Y2A * c, c, -LCAO_Y2B) : 0.f;
  Y[3] = (NL > 3) ? LCAO_Y3 * c * fmaf(5.0f * c, c, -3.0f) : 0.f;
}
template <int NL>
__device__ __forceinline__ void sph_harm_grad(float c, float (&dY)[4]) {
  dY[0] = 0.f;
  dY[1] = (NL > 1) ? LCAO_Y1 : 0.f;
  dY[2] = (NL > 2) ? 2.0f * LCAO_Y2A * c : 0.f;
  dY[3] = (NL > 3) ? LCAO_Y3 * fmaf(15.0f * c, c, -3.0f) : 0.f;
}

// |sum_l Y_l B_l|^2 = Y^T G Y from the upper-triangular FP64 Gram matrix g (NL(NL+1)/2 entries)
template <int NL>
__device__ __forceinline__ double quad_form(const double* g, const float (&Y)[4]) {
  double s = 0.0;
  int i = 0;
#pragma unroll
  for (int a = 0; a < NL; ++a)
#pragma unroll
    for (int b = a; b < NL; ++b) {
      const double t = (double)Y[a] * (double)Y[b] * g[i++];
      s += (a == b) ? t : 2.0 * t;
    }
  return s;
}
// sum_l X_l (G Y)_l  (bilinear form with the symmetric Gram matrix)
template <int NL>
__device__ __forceinline__ double bilin_form(const double* g, const float (&X)[4], const float (&Y)[4]) {
  double s = 0.0;
  int i = 0;
#pragma unroll
  for (int a = 0; a < NL; ++a)
#pragma unroll
    for (int b = a; b < NL; ++b) {
      const double gg = g[i++];
      s += (a == b) ? (double)X[a] * Y[a] * gg : ((double)X[a] * Y[b] + (double)X[b] * Y[a]) * gg;
    }
  return s;
}

__device__ __forceinline__ float4 fma4(float a, float4 x, float4 acc) {
  return make_float4(fmaf(a, x.x, acc.x), fmaf(a, x.y, acc.y), fmaf(a, x.z, acc.z), fmaf(a, x.w, acc.w));
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 scale4(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 sigmoid4(float4 x) {
  return make_float4(sigmoidf_acc(x.x), sigmoidf_acc(x.y), sigmoidf_acc(x.z), sigmoidf_acc(x.w));
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float comp4(const float4& a, int l) { return l == 0 ? a.x : l == 1 ? a.y : l == 2 ? a.z : a.w; }


}  // namespace
