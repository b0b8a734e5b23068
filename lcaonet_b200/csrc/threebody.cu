// Fused three-body message passing (lcaonet.py:173-189 + shbf.py:75-87 + lcaonet.py:431-435),
// forward and backward, one CTA per centre node s.
//
// All out-edges e=(s->t) of a node share the same set of in-edges e'=(k->s), so the per-in-edge
// data B[e',l,:] (orbital contraction grouped by angular momentum l, NL*C floats) and the gate
// sigmoid(xk[k,:]) are staged ONCE per node in shared memory and serve deg_out*deg_in triplets.
// No triplet-sized tensor (the reference's (T,O,C) gather, cos(theta), Y_l(T,O)) ever exists:
// cos(theta) = unit[e].unit[e'] and Y_l are recomputed in registers.
//
// Thread mapping: a "slot" = 8 lanes owning one out-edge (forward) / one in-edge (backward); the 8
// lanes split the C channels as interleaved float4 (lane l8 owns float4 columns l8, l8+8, ...), so
// smem reads are conflict-free 128-byte rows, the 4 slots of a warp read the same staged row
// (broadcast), and the L2 norm over channels is a 3-step xor-shuffle inside the octet.  Sums over
// triplets accumulate in registers and are written once: no atomics, deterministic.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;           // 4 warps = 16 slots
constexpr int kSlots = kThreads / 8;
constexpr float kEps = 1e-12f;          // F.normalize eps (lcaonet.py:184)

template <int NL>
__device__ __forceinline__ void sph_harm(float c, float (&Y)[NL]) {
  Y[0] = LCAO_Y0;
  if (NL > 1) Y[1] = LCAO_Y1 * c;
  if (NL > 2) Y[2] = fmaf(LCAO_Y2A * c, c, -LCAO_Y2B);
  if (NL > 3) Y[3] = LCAO_Y3 * c * fmaf(5.0f * c, c, -3.0f);
}

__device__ __forceinline__ float4 fma4(float a, float4 x, float4 acc) {
  return make_float4(fmaf(a, x.x, acc.x), fmaf(a, x.y, acc.y), fmaf(a, x.z, acc.z), fmaf(a, x.w, acc.w));
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 scale4(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ float4 sigmoid4(float4 x) {
  return make_float4(sigmoidf_acc(x.x), sigmoidf_acc(x.y), sigmoidf_acc(x.z), sigmoidf_acc(x.w));
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ---------------------------------------------------------------------------------------------
// forward: tbw[e,:] = sum_{e' in in(s), e' != e} normalize(sum_l Y_l(c) B[e',l,:]) * gate[e',:]
// ---------------------------------------------------------------------------------------------
template <int NL, int V4>
__global__ void __launch_bounds__(kThreads) k_threebody_fwd(
    const float* __restrict__ B, int NG, const float* __restrict__ unit, const float* __restrict__ xk, int64_t ldxk,
    const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_edge, const int32_t* __restrict__ in_src,
    const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_edge, int C, int TI,
    float* __restrict__ tbw) {
  extern __shared__ __align__(16) float smem[];
  const int rowC = (NL + 1) * C;
  float* tile = smem;                               // TI x (NL+1) x C : B groups then gate
  float* u_in = tile + (size_t)TI * rowC;           // TI x 3
  int* eid = reinterpret_cast<int*>(u_in + TI * 3); // TI

  const int s = blockIdx.x;
  const int ib = in_ptr[s], ie = in_ptr[s + 1], ob = out_ptr[s], oe = out_ptr[s + 1];
  if (oe == ob) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l8 = lane & 7;
  const int slot = warp * 4 + (lane >> 3);

  for (int p0 = ob; p0 < oe; p0 += 2 * kSlots) {
    int e_r[2];
    float ux[2], uy[2], uz[2];
    bool act[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int pos = p0 + r * kSlots + slot;
      act[r] = pos < oe;
      e_r[r] = act[r] ? out_edge[pos] : -1;
      ux[r] = act[r] ? unit[3 * (int64_t)e_r[r]] : 0.f;
      uy[r] = act[r] ? unit[3 * (int64_t)e_r[r] + 1] : 0.f;
      uz[r] = act[r] ? unit[3 * (int64_t)e_r[r] + 2] : 0.f;
    }
    const bool warp_r0 = __any_sync(0xffffffffu, act[0]);
    const bool warp_r1 = __any_sync(0xffffffffu, act[1]);
    float4 acc[2][V4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < V4; ++j) acc[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int c0 = ib; c0 < ie; c0 += TI) {
      const int nI = min(TI, ie - c0);
      __syncthreads();
      for (int i = warp; i < nI; i += kThreads / 32) {
        const int ep = in_edge[c0 + i];
        const int k = in_src[c0 + i];
        for (int c = lane * 4; c < C; c += 128) {
#pragma unroll
          for (int l = 0; l < NL; ++l) st4(tile + i * rowC + l * C + c, ldg4(B + ((int64_t)ep * NG + l) * C + c));
          st4(tile + i * rowC + NL * C + c, sigmoid4(ldg4(xk + (int64_t)k * ldxk + c)));
        }
        if (lane < 3) u_in[i * 3 + lane] = unit[3 * (int64_t)ep + lane];
        if (lane == 3) eid[i] = ep;
      }
      __syncthreads();
      if (!warp_r0) continue;  // warp-uniform: this warp owns no out-edge in this pass
      for (int i = 0; i < nI; ++i) {
        const float vx = u_in[3 * i], vy = u_in[3 * i + 1], vz = u_in[3 * i + 2];
        const int ep = eid[i];
        const float* row = tile + i * rowC;
        float Y[2][NL];
        sph_harm<NL>(fmaf(ux[0], vx, fmaf(uy[0], vy, uz[0] * vz)), Y[0]);
        sph_harm<NL>(fmaf(ux[1], vx, fmaf(uy[1], vy, uz[1] * vz)), Y[1]);
        float4 v[2][V4];
        float n2[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < V4; ++j) {
          const int c = (j * 8 + l8) * 4;
          if (c < C) {
            float4 b[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) b[l] = lds4(row + l * C + c);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              if (r == 1 && !warp_r1) continue;
              float4 t = scale4(Y[r][0], b[0]);
#pragma unroll
              for (int l = 1; l < NL; ++l) t = fma4(Y[r][l], b[l], t);
              v[r][j] = t;
              n2[r] += dot4(t, t);
            }
          } else {
            v[0][j] = v[1][j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float w[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (r == 1 && !warp_r1) { w[r] = 0.f; continue; }
          const float nn = octet_sum(n2[r]);
          w[r] = (act[r] && ep != e_r[r]) ? 1.0f / fmaxf(sqrtf(nn), kEps) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < V4; ++j) {
          const int c = (j * 8 + l8) * 4;
          if (c < C) {
            const float4 gate = lds4(row + NL * C + c);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              if (r == 1 && !warp_r1) continue;
              acc[r][j] = fma4(w[r], mul4(v[r][j], gate), acc[r][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (!act[r]) continue;
#pragma unroll
      for (int j = 0; j < V4; ++j) {
        const int c = (j * 8 + l8) * 4;
        if (c < C) st4(tbw + (int64_t)e_r[r] * C + c, acc[r][j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  Slot <-> in-edge e' (its B row and gate live in registers), loop over the staged
// out-edge gradients G[e,:] = d tbw[e,:].  With y = v/n, n = max(|v|, eps), sg = sigmoid(xk[k]):
//   dy = G*sg ; dv = (dy - y (y.dy)) / n  (dy/eps when clamped) ; dB[e',l,:] += Y_l dv ;
//   q[e',:] = sg (1-sg) * sum_e G*y            (d xk[k] = sum_{e' in out(k)} q[e'])
// ---------------------------------------------------------------------------------------------
template <int NL, int V4>
__global__ void __launch_bounds__(kThreads) k_threebody_bwd(
    const float* __restrict__ B, int NG, const float* __restrict__ unit, const float* __restrict__ xk, int64_t ldxk,
    const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_edge, const int32_t* __restrict__ in_src,
    const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_edge, int C, int TO,
    const float* __restrict__ d_tbw, float* __restrict__ dB, float* __restrict__ q) {
  extern __shared__ __align__(16) float smem[];
  float* G = smem;                                    // TO x C
  float* u_out = G + (size_t)TO * C;                  // TO x 3
  int* eid = reinterpret_cast<int*>(u_out + TO * 3);  // TO

  const int s = blockIdx.x;
  const int ib = in_ptr[s], ie = in_ptr[s + 1], ob = out_ptr[s], oe = out_ptr[s + 1];
  if (ie == ib) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l8 = lane & 7;
  const int slot = warp * 4 + (lane >> 3);

  for (int p0 = ib; p0 < ie; p0 += kSlots) {
    const int pos = p0 + slot;
    const bool act = pos < ie;
    const int ep = act ? in_edge[pos] : -1;
    const int k = act ? in_src[pos] : 0;
    const bool warp_act = __any_sync(0xffffffffu, act);
    float4 b[NL][V4], sg[V4], dacc[NL][V4], gy[V4];
#pragma unroll
    for (int j = 0; j < V4; ++j) {
      const int c = (j * 8 + l8) * 4;
      const bool ok = act && c < C;
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        b[l][j] = ok ? ldg4(B + ((int64_t)ep * NG + l) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        dacc[l][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      sg[j] = ok ? sigmoid4(ldg4(xk + (int64_t)k * ldxk + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      gy[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float vx = act ? unit[3 * (int64_t)ep] : 0.f, vy = act ? unit[3 * (int64_t)ep + 1] : 0.f,
                vz = act ? unit[3 * (int64_t)ep + 2] : 0.f;

    for (int c0 = ob; c0 < oe; c0 += TO) {
      const int nO = min(TO, oe - c0);
      __syncthreads();
      for (int i = warp; i < nO; i += kThreads / 32) {
        const int e = out_edge[c0 + i];
        for (int c = lane * 4; c < C; c += 128) st4(G + i * C + c, ldg4(d_tbw + (int64_t)e * C + c));
        if (lane < 3) u_out[i * 3 + lane] = unit[3 * (int64_t)e + lane];
        if (lane == 3) eid[i] = e;
      }
      __syncthreads();
      if (!warp_act) continue;
      for (int i = 0; i < nO; ++i) {
        float Y[NL];
        sph_harm<NL>(fmaf(u_out[3 * i], vx, fmaf(u_out[3 * i + 1], vy, u_out[3 * i + 2] * vz)), Y);
        const float live = (act && eid[i] != ep) ? 1.f : 0.f;
        float4 v[V4], dy[V4];
        float n2 = 0.f, vdy = 0.f;
#pragma unroll
        for (int j = 0; j < V4; ++j) {
          const int c = (j * 8 + l8) * 4;
          float4 t = scale4(Y[0], b[0][j]);
#pragma unroll
          for (int l = 1; l < NL; ++l) t = fma4(Y[l], b[l][j], t);
          v[j] = t;
          const float4 g = (c < C) ? lds4(G + i * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          dy[j] = mul4(g, sg[j]);
          n2 += dot4(t, t);
          vdy += dot4(t, dy[j]);
        }
        n2 = octet_sum(n2);
        vdy = octet_sum(vdy);
        const float nrm = sqrtf(n2);
        const float inv = live / fmaxf(nrm, kEps);          // 0 for the excluded pair (e' == e)
        const float coef = (nrm > kEps) ? vdy * inv * inv : 0.f;  // (y.dy)/n, dropped when clamped
#pragma unroll
        for (int j = 0; j < V4; ++j) {
          const int c = (j * 8 + l8) * 4;
          if (c < C) {
            // G*y = (dy/sg)*v*inv is recomputed from the staged G to avoid dividing by sg
            const float4 g = lds4(G + i * C + c);
            gy[j] = fma4(inv, mul4(g, v[j]), gy[j]);
            // dv = (dy - v*coef) * inv
            const float4 dv = scale4(inv, make_float4(fmaf(-coef, v[j].x, dy[j].x), fmaf(-coef, v[j].y, dy[j].y),
                                                      fmaf(-coef, v[j].z, dy[j].z), fmaf(-coef, v[j].w, dy[j].w)));
#pragma unroll
            for (int l = 0; l < NL; ++l) dacc[l][j] = fma4(Y[l], dv, dacc[l][j]);
          }
        }
      }
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < V4; ++j) {
        const int c = (j * 8 + l8) * 4;
        if (c < C) {
#pragma unroll
          for (int l = 0; l < NL; ++l) st4(dB + ((int64_t)ep * NG + l) * C + c, dacc[l][j]);
          for (int l = NL; l < NG; ++l) st4(dB + ((int64_t)ep * NG + l) * C + c, make_float4(0.f, 0.f, 0.f, 0.f));
          const float4 s1 = make_float4(sg[j].x * (1.f - sg[j].x), sg[j].y * (1.f - sg[j].y), sg[j].z * (1.f - sg[j].z),
                                        sg[j].w * (1.f - sg[j].w));
          st4(q + (int64_t)ep * C + c, mul4(gy[j], s1));
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

inline int v4_for(int C) { return C <= 32 ? 1 : C <= 64 ? 2 : C <= 128 ? 4 : 8; }

}  // namespace

#define TB_DISPATCH(NL, V4, CALL)                  \
  switch ((NL) * 10 + (V4)) {                      \
    case 11: { CALL(1, 1); } break;                \
    case 12: { CALL(1, 2); } break;                \
    case 14: { CALL(1, 4); } break;                \
    case 18: { CALL(1, 8); } break;                \
    case 21: { CALL(2, 1); } break;                \
    case 22: { CALL(2, 2); } break;                \
    case 24: { CALL(2, 4); } break;                \
    case 28: { CALL(2, 8); } break;                \
    case 31: { CALL(3, 1); } break;                \
    case 32: { CALL(3, 2); } break;                \
    case 34: { CALL(3, 4); } break;                \
    case 38: { CALL(3, 8); } break;                \
    case 41: { CALL(4, 1); } break;                \
    case 42: { CALL(4, 2); } break;                \
    case 44: { CALL(4, 4); } break;                \
    default: { CALL(4, 8); } break;                \
  }

extern "C" int lcao_threebody_fwd(const float* B, int32_t NG, const float* unit, const float* xk, int64_t ldxk,
                                  const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src,
                                  const int32_t* out_ptr, const int32_t* out_edge, int64_t N, int64_t E, int32_t C,
                                  int32_t NL, float* tbw, void* stream) {
  if (N == 0 || E == 0) return LCAO_OK;
  LCAO_REQUIRE(B && unit && xk && in_ptr && in_edge && in_src && out_ptr && out_edge && tbw, "lcao_threebody_fwd: null buffer");
  LCAO_REQUIRE(C % 4 == 0 && C > 0 && C <= 256 && NL >= 1 && NL <= 4 && NG >= NL && ldxk % 4 == 0,
               "lcao_threebody_fwd: need C %% 4 == 0, C <= 256, 1 <= NL <= 4, NG >= NL (C=%d NL=%d NG=%d)", C, NL, NG);
  cudaStream_t st = (cudaStream_t)stream;
  const int TI = 16;
  const size_t smem = (size_t)TI * ((NL + 1) * C + 4) * sizeof(float);
  const int V4 = v4_for(C);
#define CALL(nl, v4)                                                                                        \
  if (int rc = set_smem(k_threebody_fwd<nl, v4>, smem)) return rc;                                          \
  k_threebody_fwd<nl, v4><<<(unsigned)N, kThreads, smem, st>>>(B, NG, unit, xk, ldxk, in_ptr, in_edge, in_src, \
                                                               out_ptr, out_edge, C, TI, tbw)
  TB_DISPATCH(NL, V4, CALL)
#undef CALL
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_threebody_bwd(const float* B, int32_t NG, const float* unit, const float* xk, int64_t ldxk,
                                  const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src,
                                  const int32_t* out_ptr, const int32_t* out_edge, int64_t N, int64_t E, int32_t C,
                                  int32_t NL, const float* d_tbw, float* dB, float* q, float* d_unit_ks,
                                  float* d_unit_st, void* stream) {
  if (N == 0 || E == 0) return LCAO_OK;
  LCAO_REQUIRE(B && unit && xk && in_ptr && in_edge && in_src && out_ptr && out_edge && d_tbw && dB && q,
               "lcao_threebody_bwd: null buffer");
  LCAO_REQUIRE(C % 4 == 0 && C > 0 && C <= 256 && NL >= 1 && NL <= 4 && NG >= NL && ldxk % 4 == 0,
               "lcao_threebody_bwd: need C %% 4 == 0, C <= 256, 1 <= NL <= 4, NG >= NL");
  if (d_unit_ks || d_unit_st) {
    lcao_set_error("lcao_threebody_bwd: geometry gradients (d_unit) are not implemented yet");
    return LCAO_E_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int TO = 32;
  const size_t smem = (size_t)TO * (C + 4) * sizeof(float);
  const int V4 = v4_for(C);
#define CALL(nl, v4)                                                                                        \
  if (int rc = set_smem(k_threebody_bwd<nl, v4>, smem)) return rc;                                          \
  k_threebody_bwd<nl, v4><<<(unsigned)N, kThreads, smem, st>>>(B, NG, unit, xk, ldxk, in_ptr, in_edge, in_src, \
                                                               out_ptr, out_edge, C, TO, d_tbw, dB, q)
  TB_DISPATCH(NL, V4, CALL)
#undef CALL
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}
