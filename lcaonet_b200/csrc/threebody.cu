// Fused three-body message passing (lcaonet.py:173-189 + shbf.py:75-87 + lcaonet.py:431-435),
// forward and backward, one CTA per centre node s.
//
// All out-edges e=(s->t) of a node share the same set of in-edges e'=(k->s).  With
//   B[e',l,:]  = sum_{o in l} rb[e',o] cst'[e',o,:]        (orbital contraction grouped by l, pair_table.cu)
//   GB[e',l,:] = sigmoid(xk[k,:]) * B[e',l,:]              (gate folded in once per in-edge)
//   G[e']      = Gram matrix B[e',l,:].B[e',l',:]          (NL x NL, FP64, once per in-edge)
// the reference's per-triplet chain  v = sum_l Y_l(c) B_l ; y = v / max(|v|, eps) ; tbw[e] += y * gate
// becomes, exactly,
//   |v|^2 = Y^T G Y   (a 3x3 quadratic form per triplet instead of a C-wide reduction)
//   tbw[e,:] = sum_{e'} sum_l  a[e,e',l] * GB[e',l,:],      a = Y_l(c) / max(|v|, eps)
// i.e. per node one small dense product (deg_out x deg_in*NL) x (deg_in*NL x C) on the FP32 pipes:
// NL FMAs per triplet-channel, no cross-lane reduction in the forward at all.  No triplet-sized tensor
// (the reference's (T,O,C) gather, cos(theta), Y_l(T,O)) exists; cos(theta) = unit[e].unit[e'] and Y_l
// are recomputed per (e,e') pair by one thread and broadcast through shared memory.
//
// Thread mapping: lane <-> float4 channel column (C = 128 -> 32 lanes), warp <-> a set of out-edges
// (forward: up to 8 register accumulators) or in-edges (backward: 2 at a time).  Sums over triplets
// accumulate in registers and are written once: no atomics, deterministic.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 128;  // 4 warps
constexpr int kWarpsTb = kThreads / 32;
constexpr int kTO = 32;        // out-edges per pass (forward: 4 warps x 8 accumulators)
constexpr int kR = 8;          // forward: out-edge accumulators per warp
constexpr int kRI = 2;         // backward: in-edges per warp at a time
constexpr float kEps = 1e-12f; // F.normalize eps (lcaonet.py:184)

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

// Sum 8 per-lane partials over the 32 lanes with 9 shuffles; every lane ends up with the total of
// element  4*bit4(lane) + 2*bit3(lane) + bit2(lane).
__device__ __forceinline__ float bfly8(const float (&p)[8], int lane) {
  float q4[4], q2[2];
  bool up = (lane & 16) != 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float recv = __shfl_xor_sync(0xffffffffu, up ? p[k] : p[k + 4], 16);
    q4[k] = (up ? p[k + 4] : p[k]) + recv;
  }
  up = (lane & 8) != 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float recv = __shfl_xor_sync(0xffffffffu, up ? q4[k] : q4[k + 2], 8);
    q2[k] = (up ? q4[k + 2] : q4[k]) + recv;
  }
  up = (lane & 4) != 0;
  const float recv = __shfl_xor_sync(0xffffffffu, up ? q2[0] : q2[1], 4);
  float r = (up ? q2[1] : q2[0]) + recv;
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
__device__ __forceinline__ int bfly8_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int NL, int V4>
__global__ void __launch_bounds__(kThreads) k_threebody_fwd(
    const float* __restrict__ B, int NG, const double* __restrict__ gram, const float* __restrict__ unit,
    const float* __restrict__ xk, int64_t ldxk, const int32_t* __restrict__ in_ptr,
    const int32_t* __restrict__ in_edge, const int32_t* __restrict__ in_src, const int32_t* __restrict__ out_ptr,
    const int32_t* __restrict__ out_edge, int C, int TI, float* __restrict__ tbw) {
  constexpr int NP = NL * (NL + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* g_in = reinterpret_cast<double*>(smem_raw);                 // TI x NP
  float* GB = reinterpret_cast<float*>(g_in + (size_t)TI * NP);      // TI x NL x C
  float* A = GB + (size_t)TI * NL * C;                               // kTO x TI x 4
  float* u_in = A + (size_t)kTO * TI * 4;                            // TI x 3
  float* u_out = u_in + TI * 3;                                      // kTO x 3
  int* eid_in = reinterpret_cast<int*>(u_out + kTO * 3);             // TI
  int* eid_out = eid_in + TI;                                        // kTO

  const int s = blockIdx.x;
  const int ib = in_ptr[s], ie = in_ptr[s + 1], ob = out_ptr[s], oe = out_ptr[s + 1];
  if (oe == ob) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int p0 = ob; p0 < oe; p0 += kTO) {
    const int nO = min(kTO, oe - p0);
    __syncthreads();  // previous pass has finished reading u_out / eid_out
    if (threadIdx.x < nO) {
      const int e = out_edge[p0 + threadIdx.x];
      eid_out[threadIdx.x] = e;
      u_out[3 * threadIdx.x] = unit[3 * (int64_t)e];
      u_out[3 * threadIdx.x + 1] = unit[3 * (int64_t)e + 1];
      u_out[3 * threadIdx.x + 2] = unit[3 * (int64_t)e + 2];
    }
    __syncthreads();
    const int nr = (nO > warp) ? (nO - warp + kWarpsTb - 1) / kWarpsTb : 0;  // this warp's out-edges: j = r*4 + warp
    float4 acc[kR][V4];
#pragma unroll
    for (int r = 0; r < kR; ++r)
#pragma unroll
      for (int v = 0; v < V4; ++v) acc[r][v] = zero4();

    for (int c0 = ib; c0 < ie; c0 += TI) {
      const int nI = min(TI, ie - c0);
      __syncthreads();  // previous chunk's consumers are done with GB / A
      // ---- stage the in-edges of this chunk: gated B rows, directions, ids, Gram matrices
      for (int i = warp; i < nI; i += kWarpsTb) {
        const int ep = in_edge[c0 + i];
        const int k = in_src[c0 + i];
#pragma unroll
        for (int v = 0; v < V4; ++v) {
          const int c = (lane + 32 * v) * 4;
          if (c < C) {
            const float4 gate = sigmoid4(ldg4(xk + (int64_t)k * ldxk + c));
#pragma unroll
            for (int l = 0; l < NL; ++l)
              st4(GB + ((size_t)i * NL + l) * C + c, mul4(gate, ldg4(B + ((int64_t)ep * NG + l) * C + c)));
          }
        }
        if (lane < 3) u_in[i * 3 + lane] = unit[3 * (int64_t)ep + lane];
        if (lane == 3) eid_in[i] = ep;
        if (lane >= 4 && lane < 4 + NP) g_in[i * NP + lane - 4] = gram[(int64_t)ep * NP + lane - 4];
      }
      __syncthreads();
      // ---- per (out-edge j, in-edge i): a_l = Y_l(c) / max(|v|, eps), 0 for the excluded pair e' == e
      for (int t = threadIdx.x; t < nO * nI; t += kThreads) {
        const int j = t / nI, i = t - j * nI;
        const float c = fmaf(u_out[3 * j], u_in[3 * i], fmaf(u_out[3 * j + 1], u_in[3 * i + 1], u_out[3 * j + 2] * u_in[3 * i + 2]));
        float Y[4];
        sph_harm<NL>(c, Y);
        const float n2 = fmaxf((float)quad_form<NL>(g_in + i * NP, Y), 0.f);
        const float w = (eid_out[j] != eid_in[i]) ? 1.0f / fmaxf(sqrtf(n2), kEps) : 0.f;
        st4(A + ((size_t)j * TI + i) * 4, make_float4(w * Y[0], w * Y[1], w * Y[2], w * Y[3]));
      }
      __syncthreads();
      // ---- tbw[j,:] += sum_i sum_l a[j,i,l] GB[i,l,:]
      if (nr > 0) {
        for (int i = 0; i < nI; ++i) {
          float4 gb[NL][V4];
#pragma unroll
          for (int v = 0; v < V4; ++v) {
            const int c = (lane + 32 * v) * 4;
#pragma unroll
            for (int l = 0; l < NL; ++l) gb[l][v] = (c < C) ? lds4(GB + ((size_t)i * NL + l) * C + c) : zero4();
          }
#pragma unroll
          for (int r = 0; r < kR; ++r) {
            if (r < nr) {
              const float4 a = lds4(A + ((size_t)(r * kWarpsTb + warp) * TI + i) * 4);
#pragma unroll
              for (int v = 0; v < V4; ++v) {
                acc[r][v] = fma4(a.x, gb[0][v], acc[r][v]);
                if (NL > 1) acc[r][v] = fma4(a.y, gb[1][v], acc[r][v]);
                if (NL > 2) acc[r][v] = fma4(a.z, gb[2][v], acc[r][v]);
                if (NL > 3) acc[r][v] = fma4(a.w, gb[3][v], acc[r][v]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      if (r < nr) {
        const int e = eid_out[r * kWarpsTb + warp];
#pragma unroll
        for (int v = 0; v < V4; ++v) {
          const int c = (lane + 32 * v) * 4;
          if (c < C) st4(tbw + (int64_t)e * C + c, acc[r][v]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  Given Gt[e,:] = d tbw[e,:], per pair (e = out j, e' = in i) with a_l = w Y_l, w = 1/max(|v|,eps):
//   dGB[i,l,:] += a_l Gt[j,:]                                   (gate and B gradients follow per in-edge)
//   dot = Gt[j,:] . sum_l a_l GB[i,l,:]  (= w * dL/dw)           one C-wide dot per pair, butterfly-reduced
//   H_i[l,l'] -= dot a_l a_l'            (norm path; skipped when |v| <= eps)  ->  dB[i,l] += sum_l' H[l,l'] B[i,l']
//   dB[i,l,:] += gate * dGB[i,l,:] ;  q[i,:] = gate (1-gate) sum_l B[i,l,:] dGB[i,l,:]   (d xk[k] = sum_{e' in out(k)} q[e'])
//   FORCES: dc = Gt[j,:] . sum_l (w Y'_l) GB[i,l,:] - w^2 dot sum_l Y'_l (G Y)_l ;  d unit[e_j] += dc unit[e_i] and v.v.
// ---------------------------------------------------------------------------------------------
template <int NL, int V4, bool FORCES>
__global__ void __launch_bounds__(kThreads) k_threebody_bwd(
    const float* __restrict__ B, int NG, const double* __restrict__ gram, const float* __restrict__ unit,
    const float* __restrict__ xk, int64_t ldxk, const int32_t* __restrict__ in_ptr,
    const int32_t* __restrict__ in_edge, const int32_t* __restrict__ in_src, const int32_t* __restrict__ out_ptr,
    const int32_t* __restrict__ out_edge, int C, int TI, const float* __restrict__ d_tbw, float* __restrict__ dB,
    float* __restrict__ q, float* __restrict__ du_ks, float* __restrict__ du_st) {
  constexpr int NP = NL * (NL + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* g_in = reinterpret_cast<double*>(smem_raw);              // TI x NP
  float* Gt = reinterpret_cast<float*>(g_in + (size_t)TI * NP);   // kTO x C
  float* A = Gt + (size_t)kTO * C;                                // kTO x TI x 4   a_l = w Y_l
  float* Fl = A + (size_t)kTO * TI * 4;                           // kTO x TI       1 if the norm path is live
  float* A2 = Fl + (size_t)kTO * TI;                              // FORCES: kTO x TI x 4   w Y'_l
  float* Cc = A2 + (FORCES ? (size_t)kTO * TI * 4 : 0);           // FORCES: kTO x TI       cos, then dL/dcos
  float* Ww = Cc + (FORCES ? (size_t)kTO * TI : 0);               // FORCES: kTO x TI       w
  float* u_in = Ww + (FORCES ? (size_t)kTO * TI : 0);             // TI x 3
  float* u_out = u_in + TI * 3;                                   // kTO x 3
  int* eid_in = reinterpret_cast<int*>(u_out + kTO * 3);          // TI
  int* eid_out = eid_in + TI;                                     // kTO

  const int s = blockIdx.x;
  const int ib = in_ptr[s], ie = in_ptr[s + 1], ob = out_ptr[s], oe = out_ptr[s + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ie == ib) {  // no in-edges: only the s->t role gradients of the out-edges exist, and they are zero
    if (FORCES)
      for (int t = threadIdx.x; t < (oe - ob) * 3; t += kThreads) du_st[3 * (int64_t)out_edge[ob + t / 3] + t % 3] = 0.f;
    return;
  }
  if (oe == ob) {  // in-edges that feed no triplet: zero gradients
    for (int i = warp; i < ie - ib; i += kWarpsTb) {
      const int ep = in_edge[ib + i];
      for (int c = lane * 4; c < C; c += 128) {
        for (int l = 0; l < NG; ++l) st4(dB + ((int64_t)ep * NG + l) * C + c, zero4());
        st4(q + (int64_t)ep * C + c, zero4());
      }
      if (FORCES && lane < 3) du_ks[3 * (int64_t)ep + lane] = 0.f;
    }
    return;
  }

  for (int p0 = ob; p0 < oe; p0 += kTO) {
    const int nO = min(kTO, oe - p0);
    const bool first_out = (p0 == ob);
    __syncthreads();
    for (int j = warp; j < nO; j += kWarpsTb) {
      const int e = out_edge[p0 + j];
      for (int c = lane * 4; c < C; c += 128) st4(Gt + (size_t)j * C + c, ldg4(d_tbw + (int64_t)e * C + c));
      if (lane < 3) u_out[j * 3 + lane] = unit[3 * (int64_t)e + lane];
      if (lane == 3) eid_out[j] = e;
    }
    for (int c0 = ib; c0 < ie; c0 += TI) {
      const int nI = min(TI, ie - c0);
      const bool first_in = (c0 == ib);
      __syncthreads();
      for (int t = threadIdx.x; t < nI; t += kThreads) {
        const int ep = in_edge[c0 + t];
        eid_in[t] = ep;
        u_in[3 * t] = unit[3 * (int64_t)ep];
        u_in[3 * t + 1] = unit[3 * (int64_t)ep + 1];
        u_in[3 * t + 2] = unit[3 * (int64_t)ep + 2];
#pragma unroll
        for (int p = 0; p < NP; ++p) g_in[t * NP + p] = gram[(int64_t)ep * NP + p];
      }
      __syncthreads();
      for (int t = threadIdx.x; t < nO * nI; t += kThreads) {
        const int j = t / nI, i = t - j * nI;
        const float c = fmaf(u_out[3 * j], u_in[3 * i], fmaf(u_out[3 * j + 1], u_in[3 * i + 1], u_out[3 * j + 2] * u_in[3 * i + 2]));
        float Y[4];
        sph_harm<NL>(c, Y);
        const float nrm = sqrtf(fmaxf((float)quad_form<NL>(g_in + i * NP, Y), 0.f));
        const bool live = eid_out[j] != eid_in[i];
        const float w = live ? 1.0f / fmaxf(nrm, kEps) : 0.f;
        const size_t o = (size_t)j * TI + i;
        st4(A + o * 4, make_float4(w * Y[0], w * Y[1], w * Y[2], w * Y[3]));
        Fl[o] = (live && nrm > kEps) ? 1.f : 0.f;
        if (FORCES) {
          float dY[4];
          sph_harm_grad<NL>(c, dY);
          st4(A2 + o * 4, make_float4(w * dY[0], w * dY[1], w * dY[2], w * dY[3]));
          Cc[o] = c;
          Ww[o] = w;
        }
      }
      __syncthreads();
      // ---- each warp owns in-edges i0 .. i0+kRI-1 of the chunk
      for (int i0 = warp * kRI; i0 < nI; i0 += kWarpsTb * kRI) {
        float4 gb[kRI][NL][V4], dacc[kRI][NL][V4], gate[kRI][V4];
        float h[kRI][NP];
        int ep[kRI];
#pragma unroll
        for (int rr = 0; rr < kRI; ++rr) {
          const bool valid = i0 + rr < nI;
          ep[rr] = valid ? eid_in[i0 + rr] : -1;
          const int k = valid ? in_src[c0 + i0 + rr] : 0;
#pragma unroll
          for (int v = 0; v < V4; ++v) {
            const int c = (lane + 32 * v) * 4;
            const bool ok = valid && c < C;
            gate[rr][v] = ok ? sigmoid4(ldg4(xk + (int64_t)k * ldxk + c)) : zero4();
#pragma unroll
            for (int l = 0; l < NL; ++l) {
              gb[rr][l][v] = ok ? mul4(gate[rr][v], ldg4(B + ((int64_t)ep[rr] * NG + l) * C + c)) : zero4();
              dacc[rr][l][v] = zero4();
            }
          }
#pragma unroll
          for (int p = 0; p < NP; ++p) h[rr][p] = 0.f;
        }
        const int jsub = bfly8_index(lane);
        for (int j0 = 0; j0 < nO; j0 += 8) {
          float part[kRI][8], part2[FORCES ? kRI : 1][8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
            for (int rr = 0; rr < kRI; ++rr) {
              part[rr][jj] = 0.f;
              if constexpr (FORCES) part2[rr][jj] = 0.f;
            }
            if (j0 + jj < nO) {  // warp-uniform
              const int j = j0 + jj;
              float4 g[V4];
#pragma unroll
              for (int v = 0; v < V4; ++v) {
                const int c = (lane + 32 * v) * 4;
                g[v] = (c < C) ? lds4(Gt + (size_t)j * C + c) : zero4();
              }
#pragma unroll
              for (int rr = 0; rr < kRI; ++rr) {
                if (i0 + rr < nI) {  // warp-uniform
                  const float4 a = lds4(A + ((size_t)j * TI + i0 + rr) * 4);
                  float4 a2;
                  if constexpr (FORCES) a2 = lds4(A2 + ((size_t)j * TI + i0 + rr) * 4);
                  float d = 0.f, d2 = 0.f;
#pragma unroll
                  for (int v = 0; v < V4; ++v) {
                    float4 t = scale4(a.x, gb[rr][0][v]);
                    if (NL > 1) t = fma4(a.y, gb[rr][1][v], t);
                    if (NL > 2) t = fma4(a.z, gb[rr][2][v], t);
                    if (NL > 3) t = fma4(a.w, gb[rr][3][v], t);
                    d += dot4(t, g[v]);
                    if constexpr (FORCES) {
                      float4 t2 = scale4(a2.x, gb[rr][0][v]);
                      if (NL > 1) t2 = fma4(a2.y, gb[rr][1][v], t2);
                      if (NL > 2) t2 = fma4(a2.z, gb[rr][2][v], t2);
                      if (NL > 3) t2 = fma4(a2.w, gb[rr][3][v], t2);
                      d2 += dot4(t2, g[v]);
                    }
                    dacc[rr][0][v] = fma4(a.x, g[v], dacc[rr][0][v]);
                    if (NL > 1) dacc[rr][1][v] = fma4(a.y, g[v], dacc[rr][1][v]);
                    if (NL > 2) dacc[rr][2][v] = fma4(a.z, g[v], dacc[rr][2][v]);
                    if (NL > 3) dacc[rr][3][v] = fma4(a.w, g[v], dacc[rr][3][v]);
                  }
                  part[rr][jj] = d;
                  if constexpr (FORCES) part2[rr][jj] = d2;
                }
              }
            }
          }
#pragma unroll
          for (int rr = 0; rr < kRI; ++rr) {
            const float dot = bfly8(part[rr], lane);
            float dot2 = 0.f;
            if constexpr (FORCES) dot2 = bfly8(part2[rr], lane);
            const int j = j0 + jsub;
            if (j < nO && i0 + rr < nI) {
              const size_t o = (size_t)j * TI + i0 + rr;
              const float4 a = lds4(A + o * 4);
              const float sc = -Fl[o] * dot;
              int p = 0;
#pragma unroll
              for (int x = 0; x < NL; ++x)
#pragma unroll
                for (int y = x; y < NL; ++y) h[rr][p++] += sc * comp4(a, x) * comp4(a, y);
              if (FORCES && (lane & 3) == 0) {
                const float c = Cc[o], w = Ww[o];
                float Y[4], dY[4];
                sph_harm<NL>(c, Y);
                sph_harm_grad<NL>(c, dY);
                const float corr = (float)bilin_form<NL>(g_in + (i0 + rr) * NP, dY, Y);
                Cc[o] = dot2 - Fl[o] * w * w * dot * corr;  // dL/dcos of this pair
              }
            }
          }
        }
        // ---- finish the in-edges of this warp
#pragma unroll
        for (int rr = 0; rr < kRI; ++rr) {
          if (i0 + rr >= nI) continue;
#pragma unroll
          for (int p = 0; p < NP; ++p) {  // h was accumulated by one lane quad per out-edge slot: add the 8 slots
            float x = h[rr][p];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            h[rr][p] = x;
          }
          float H[NL][NL];
          {
            int p = 0;
#pragma unroll
            for (int x = 0; x < NL; ++x)
#pragma unroll
              for (int y = x; y < NL; ++y) { H[x][y] = h[rr][p]; H[y][x] = h[rr][p]; ++p; }
          }
#pragma unroll
          for (int v = 0; v < V4; ++v) {
            const int c = (lane + 32 * v) * 4;
            if (c < C) {
              float4 b[NL];
#pragma unroll
              for (int l = 0; l < NL; ++l) b[l] = ldg4(B + ((int64_t)ep[rr] * NG + l) * C + c);
              float4 qq = zero4();
#pragma unroll
              for (int l = 0; l < NL; ++l) {
                float4 o4 = mul4(gate[rr][v], dacc[rr][l][v]);
#pragma unroll
                for (int l2 = 0; l2 < NL; ++l2) o4 = fma4(H[l][l2], b[l2], o4);
                float* dst = dB + ((int64_t)ep[rr] * NG + l) * C + c;
                if (!first_out) o4 = add4(o4, *reinterpret_cast<const float4*>(dst));
                st4(dst, o4);
                qq = add4(qq, mul4(b[l], dacc[rr][l][v]));
              }
              const float4 sg = gate[rr][v];
              qq = mul4(qq, make_float4(sg.x * (1.f - sg.x), sg.y * (1.f - sg.y), sg.z * (1.f - sg.z), sg.w * (1.f - sg.w)));
              float* qd = q + (int64_t)ep[rr] * C + c;
              if (!first_out) qq = add4(qq, *reinterpret_cast<const float4*>(qd));
              st4(qd, qq);
              if (first_out)
                for (int l = NL; l < NG; ++l) st4(dB + ((int64_t)ep[rr] * NG + l) * C + c, zero4());
            }
          }
        }
      }
      if (FORCES) {
        __syncthreads();  // all dL/dcos of this (out chunk, in chunk) are in Cc
        if (threadIdx.x < nO) {
          const int j = threadIdx.x;
          float ax = 0.f, ay = 0.f, az = 0.f;
          for (int i = 0; i < nI; ++i) {
            const float dc = Cc[(size_t)j * TI + i];
            ax = fmaf(dc, u_in[3 * i], ax); ay = fmaf(dc, u_in[3 * i + 1], ay); az = fmaf(dc, u_in[3 * i + 2], az);
          }
          float* d = du_st + 3 * (int64_t)eid_out[j];
          if (!first_in) { ax += d[0]; ay += d[1]; az += d[2]; }
          d[0] = ax; d[1] = ay; d[2] = az;
        } else if (threadIdx.x >= 64 && threadIdx.x - 64 < nI) {
          const int i = threadIdx.x - 64;
          float ax = 0.f, ay = 0.f, az = 0.f;
          for (int j = 0; j < nO; ++j) {
            const float dc = Cc[(size_t)j * TI + i];
            ax = fmaf(dc, u_out[3 * j], ax); ay = fmaf(dc, u_out[3 * j + 1], ay); az = fmaf(dc, u_out[3 * j + 2], az);
          }
          float* d = du_ks + 3 * (int64_t)eid_in[i];
          if (!first_out) { ax += d[0]; ay += d[1]; az += d[2]; }
          d[0] = ax; d[1] = ay; d[2] = az;
        }
      }
    }
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      lcao_set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
      return LCAO_E_CUDA;
    }
  }
  return LCAO_OK;
}

// in-edges staged per chunk; LCAO_TB_TI overrides the default for tuning runs (multiple of 4, <= 64)
int tile_in(const char* env, int dflt) {
  const char* s = getenv(env);
  int v = s ? atoi(s) : dflt;
  if (v < 4 || v > 64 || v % 4) v = dflt;
  return v;
}

}  // namespace

#define TB_DISPATCH(NL, V4, CALL)   \
  switch ((NL) * 10 + (V4)) {       \
    case 11: { CALL(1, 1); } break; \
    case 12: { CALL(1, 2); } break; \
    case 21: { CALL(2, 1); } break; \
    case 22: { CALL(2, 2); } break; \
    case 31: { CALL(3, 1); } break; \
    case 32: { CALL(3, 2); } break; \
    case 41: { CALL(4, 1); } break; \
    default: { CALL(4, 2); } break; \
  }

extern "C" int lcao_threebody_fwd(const float* B, int32_t NG, const double* gram, const float* unit, const float* xk,
                                  int64_t ldxk, const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src,
                                  const int32_t* out_ptr, const int32_t* out_edge, int64_t N, int64_t E, int32_t C,
                                  int32_t NL, float* tbw, void* stream) {
  if (N == 0 || E == 0) return LCAO_OK;
  LCAO_REQUIRE(B && gram && unit && xk && in_ptr && in_edge && in_src && out_ptr && out_edge && tbw,
               "lcao_threebody_fwd: null buffer");
  LCAO_REQUIRE(C % 4 == 0 && C > 0 && C <= 256 && NL >= 1 && NL <= 4 && NG >= NL && ldxk % 4 == 0,
               "lcao_threebody_fwd: need C %% 4 == 0, C <= 256, 1 <= NL <= 4, NG >= NL (C=%d NL=%d NG=%d)", C, NL, NG);
  cudaStream_t st = (cudaStream_t)stream;
  static const int TI = tile_in("LCAO_TB_TI", 32);
  const int NP = NL * (NL + 1) / 2;
  const size_t smem = (size_t)TI * NP * 8 + sizeof(float) * ((size_t)TI * NL * C + (size_t)kTO * TI * 4 + TI * 3 + kTO * 3 + TI + kTO);
  const int V4 = C <= 128 ? 1 : 2;
#define CALL(nl, v4)                                                                                             \
  if (int rc = set_smem(k_threebody_fwd<nl, v4>, smem)) return rc;                                               \
  k_threebody_fwd<nl, v4><<<(unsigned)N, kThreads, smem, st>>>(B, NG, gram, unit, xk, ldxk, in_ptr, in_edge, in_src, \
                                                               out_ptr, out_edge, C, TI, tbw)
  TB_DISPATCH(NL, V4, CALL)
#undef CALL
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_threebody_bwd(const float* B, int32_t NG, const double* gram, const float* unit, const float* xk,
                                  int64_t ldxk, const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src,
                                  const int32_t* out_ptr, const int32_t* out_edge, int64_t N, int64_t E, int32_t C,
                                  int32_t NL, const float* d_tbw, float* dB, float* q, float* d_unit_ks,
                                  float* d_unit_st, void* stream) {
  if (N == 0 || E == 0) return LCAO_OK;
  LCAO_REQUIRE(B && gram && unit && xk && in_ptr && in_edge && in_src && out_ptr && out_edge && d_tbw && dB && q,
               "lcao_threebody_bwd: null buffer");
  LCAO_REQUIRE(C % 4 == 0 && C > 0 && C <= 256 && NL >= 1 && NL <= 4 && NG >= NL && ldxk % 4 == 0,
               "lcao_threebody_bwd: need C %% 4 == 0, C <= 256, 1 <= NL <= 4, NG >= NL");
  LCAO_REQUIRE((d_unit_ks == nullptr) == (d_unit_st == nullptr), "lcao_threebody_bwd: pass both d_unit buffers or neither");
  cudaStream_t st = (cudaStream_t)stream;
  static const int TI = tile_in("LCAO_TB_TI_BWD", 32);
  const int NP = NL * (NL + 1) / 2;
  const bool forces = d_unit_ks != nullptr;
  const size_t pairs = (size_t)kTO * TI;
  const size_t smem = (size_t)TI * NP * 8 +
                      sizeof(float) * ((size_t)kTO * C + pairs * 4 + pairs + (forces ? pairs * 6 : 0) + TI * 3 + kTO * 3 + TI + kTO);
  const int V4 = C <= 128 ? 1 : 2;
#define CALL(nl, v4)                                                                                                  \
  if (forces) {                                                                                                       \
    if (int rc = set_smem(k_threebody_bwd<nl, v4, true>, smem)) return rc;                                            \
    k_threebody_bwd<nl, v4, true><<<(unsigned)N, kThreads, smem, st>>>(B, NG, gram, unit, xk, ldxk, in_ptr, in_edge,  \
                                                                       in_src, out_ptr, out_edge, C, TI, d_tbw, dB, q, \
                                                                       d_unit_ks, d_unit_st);                         \
  } else {                                                                                                            \
    if (int rc = set_smem(k_threebody_bwd<nl, v4, false>, smem)) return rc;                                           \
    k_threebody_bwd<nl, v4, false><<<(unsigned)N, kThreads, smem, st>>>(B, NG, gram, unit, xk, ldxk, in_ptr, in_edge, \
                                                                        in_src, out_ptr, out_edge, C, TI, d_tbw, dB,  \
                                                                        q, nullptr, nullptr);                         \
  }
  TB_DISPATCH(NL, V4, CALL)
#undef CALL
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}
