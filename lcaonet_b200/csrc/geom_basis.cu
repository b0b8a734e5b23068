// Edge geometry + hydrogen-like radial basis with smooth cutoff, forward and position-backward.
// Reference: base.py:27-43 (vectors under PBC), cutoff.py:32-67, rbf.py:92-103,129-142,164-182.
// One thread per edge; arithmetic in fp64 (E-sized, negligible cost) so that the fp32 outputs are
// correctly rounded images of the fp64 oracle: the cutoff is evaluated in its factored form
// fc = (1-q)^3 (1+3q+6q^2) which has no cancellation at the cutoff shell.
#include "common.cuh"

namespace {

struct CutVal { double f, df; };

__device__ __forceinline__ CutVal cutoff_eval(int kind, double r, double rc) {
  CutVal c{0.0, 0.0};
  if (!(r <= rc)) return c;
  const double q = r / rc, u = 1.0 - q;
  if (kind == LCAO_CUT_POLYNOMIAL) {
    c.f = u * u * u * (1.0 + q * (3.0 + 6.0 * q));
    c.df = -30.0 * q * q * u * u / rc;
  } else if (kind == LCAO_CUT_ENVELOPE) {  // p = 5
    c.f = u * u * u * (1.0 + q * (3.0 + q * (6.0 + q * (10.0 + 15.0 * q))));
    c.df = -105.0 * q * q * q * q * u * u / rc;
  } else {  // cosine
    const double a = 3.14159265358979323846 / rc;
    c.f = 0.5 * (cos(a * r) + 1.0);
    c.df = -0.5 * a * sin(a * r);
  }
  return c;
}

__global__ void __launch_bounds__(128) k_geom_basis_fwd(
    const float* __restrict__ pos, const float* __restrict__ shift, const float* __restrict__ lattice,
    const int64_t* __restrict__ batch, const int32_t* __restrict__ src32, const int32_t* __restrict__ dst32,
    int64_t E, const __grid_constant__ lcao_basis_spec sp, float* __restrict__ dist, float* __restrict__ unit,
    float* __restrict__ rb, float* __restrict__ drb) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t s = src32[e], t = dst32[e];
  const float* L = lattice + 9 * (batch ? batch[s] : 0);
  const double s0 = shift[3 * e], s1 = shift[3 * e + 1], s2 = shift[3 * e + 2];
  double v[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    v[j] = ((double)pos[3 * t + j] - (double)pos[3 * s + j]) + (s0 * L[j] + s1 * L[3 + j] + s2 * L[6 + j]);
  const double r = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  dist[e] = (float)r;
#pragma unroll
  for (int j = 0; j < 3; ++j) unit[3 * e + j] = (float)(v[j] / r);
  const CutVal fc = cutoff_eval(sp.cutoff_kind, r, sp.rc);
  const int O = sp.n_unique * sp.n_rep;
  for (int u = 0; u < sp.n_unique; ++u) {
    double R, dR;
    if (sp.rbf_kind == LCAO_RBF_HYDROGEN) {
      const double zs = 2.0 / (sp.n[u] * sp.a0), zeta = zs * r;
      double p = 0.0, dp = 0.0;
      for (int i = sp.deg[u]; i >= 0; --i) {  // Horner for value and derivative
        dp = dp * zeta + p;
        p = p * zeta + sp.poly[u][i];
      }
      const int l = sp.l[u];
      double zl = 1.0, zlm1 = 0.0;  // zeta^l and l*zeta^(l-1)
      for (int i = 0; i < l; ++i) { zlm1 = zl * (i + 1); zl *= zeta; }
      if (l == 0) zlm1 = 0.0;
      const double ex = exp(-0.5 * zeta);
      R = sp.norm[u] * p * zl * ex;
      dR = sp.norm[u] * ex * (dp * zl + p * zlm1 - 0.5 * p * zl) * zs;
    } else {  // spherical Bessel j0-like: sin(pi n r / rc) / r
      const double w = 3.14159265358979323846 * sp.n[u] / sp.rc;
      R = sin(w * r) / r;
      dR = w * cos(w * r) / r - sin(w * r) / (r * r);
    }
    const float val = (float)(fc.f * R), dval = (float)(fc.df * R + fc.f * dR);
    for (int k = 0; k < sp.n_rep; ++k) {
      rb[e * O + u * sp.n_rep + k] = val;
      if (drb) drb[e * O + u * sp.n_rep + k] = dval;
    }
  }
}

// dvec[e] = (d_unit - unit (unit . d_unit)) / r + (d_dist + sum_o d_rb * drb) unit
__global__ void k_geom_edge_bwd(const float* __restrict__ dist, const float* __restrict__ unit,
                                const float* __restrict__ drb, const float* __restrict__ d_dist,
                                const float* __restrict__ d_unit, const float* __restrict__ d_rb, int64_t E, int O,
                                float* __restrict__ dvec) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float ux = unit[3 * e], uy = unit[3 * e + 1], uz = unit[3 * e + 2];
  float gr = d_dist ? d_dist[e] : 0.f;
  if (d_rb)
    for (int o = 0; o < O; ++o) gr += d_rb[e * O + o] * drb[e * O + o];
  float gx = gr * ux, gy = gr * uy, gz = gr * uz;
  if (d_unit) {
    const float ax = d_unit[3 * e], ay = d_unit[3 * e + 1], az = d_unit[3 * e + 2];
    const float dot = ax * ux + ay * uy + az * uz, inv = 1.0f / dist[e];
    gx += (ax - ux * dot) * inv;
    gy += (ay - uy * dot) * inv;
    gz += (az - uz * dot) * inv;
  }
  dvec[3 * e] = gx; dvec[3 * e + 1] = gy; dvec[3 * e + 2] = gz;
}

// d_pos[n] = sum_{e in in(n)} dvec[e] - sum_{e in out(n)} dvec[e]   (vec = pos[t] - pos[s] + ...)
__global__ void k_geom_node_bwd(const float* __restrict__ dvec, const int32_t* __restrict__ in_ptr,
                                const int32_t* __restrict__ in_edge, const int32_t* __restrict__ out_ptr,
                                const int32_t* __restrict__ out_edge, int64_t N, float* __restrict__ d_pos) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= N * 3) return;
  const int64_t n = idx / 3;
  const int j = (int)(idx - n * 3);
  float acc = 0.f;
  for (int32_t p = in_ptr[n]; p < in_ptr[n + 1]; ++p) acc += dvec[3 * (int64_t)in_edge[p] + j];
  for (int32_t p = out_ptr[n]; p < out_ptr[n + 1]; ++p) acc -= dvec[3 * (int64_t)out_edge[p] + j];
  d_pos[idx] = acc;
}

}  // namespace

extern "C" int lcao_geom_basis_fwd(const float* pos, const float* shift, const float* lattice, const int64_t* batch,
                                   const int32_t* src32, const int32_t* dst32, int64_t E,
                                   const lcao_basis_spec* sp, float* dist, float* unit, float* rb, float* drb,
                                   void* stream) {
  if (E == 0) return LCAO_OK;
  LCAO_REQUIRE(pos && shift && lattice && src32 && dst32 && sp && dist && unit && rb, "lcao_geom_basis_fwd: null buffer");
  LCAO_REQUIRE(sp->n_unique >= 1 && sp->n_unique <= LCAO_MAX_UNIQUE_ORB && sp->n_rep >= 1,
               "lcao_geom_basis_fwd: bad orbital count");
  for (int u = 0; u < sp->n_unique; ++u)
    LCAO_REQUIRE(sp->deg[u] >= 0 && sp->deg[u] < LCAO_MAX_POLY && sp->l[u] >= 0 && sp->l[u] <= 3,
                 "lcao_geom_basis_fwd: bad orbital entry %d", u);
  k_geom_basis_fwd<<<(unsigned)ceil_div64(E, 128), 128, 0, (cudaStream_t)stream>>>(pos, shift, lattice, batch, src32,
                                                                                  dst32, E, *sp, dist, unit, rb, drb);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_geom_basis_bwd(const float* dist, const float* unit, const float* drb, const float* d_dist,
                                   const float* d_unit, const float* d_rb, int64_t E, int64_t N, int32_t O,
                                   const int32_t* in_ptr, const int32_t* in_edge, const int32_t* out_ptr,
                                   const int32_t* out_edge, float* dvec, float* d_pos, void* stream) {
  LCAO_REQUIRE(d_pos && in_ptr && out_ptr, "lcao_geom_basis_bwd: null buffer");
  LCAO_REQUIRE(!d_rb || drb, "lcao_geom_basis_bwd: d_rb needs the saved drb");
  cudaStream_t st = (cudaStream_t)stream;
  if (E > 0) {
    LCAO_REQUIRE(dist && unit && dvec && in_edge && out_edge, "lcao_geom_basis_bwd: null edge buffer");
    k_geom_edge_bwd<<<(unsigned)ceil_div64(E, 256), 256, 0, st>>>(dist, unit, drb, d_dist, d_unit, d_rb, E, O, dvec);
    LCAO_LAUNCH_CHECK();
  }
  if (N > 0) {
    k_geom_node_bwd<<<(unsigned)ceil_div64(N * 3, 256), 256, 0, st>>>(dvec, in_ptr, in_edge, out_ptr, out_edge, N, d_pos);
    LCAO_LAUNCH_CHECK();
  }
  return LCAO_OK;
}
