// Count-weighted BatchNorm over table rows, forward and backward (embed.py:175,194,232,249 = nn.BatchNorm1d).
//
// The reference normalises (N, F) node rows and (E, O*K) coefficient rows; both are functions of the species (pair)
// only, so the batch is given as its R DISTINCT rows plus their multiplicities: with n = sum_r c_r and
// w_r = c_r / n,
//   mean = sum_r w_r x_r ;  var = sum_r w_r (x_r - mean)^2 (biased) ;  y_r = (x_r - mean) rstd * gamma + beta
//   running_mean <- (1-m) running_mean + m mean ;  running_var <- (1-m) running_var + m var n/(n-1) ;  tracked += 1
// and, for the backward pass (rows with c_r = 0 still receive y_r, hence the un-weighted sums over dy):
//   dbeta = sum_r dy_r ;  dgamma = sum_r dy_r xhat_r ;  dx_r = rstd gamma (dy_r - w_r dbeta - w_r xhat_r dgamma)
// (eval mode: mean / var are the running statistics and dx_r = rstd gamma dy_r).
//
// CTA = 32 feature columns x 32 row lanes (row lane j takes rows j, j+32, ...); column sums are combined through
// shared memory in a fixed order: deterministic.  The tables have 37 .. (max_z+1)^2 rows of 128 .. 1024 features
// (5.6 MB at most, L2 resident): the job of these two kernels is to replace ~100 torch launches per step, not bandwidth.
#include "common.cuh"

namespace {

constexpr int kCols = 32, kLanes = 32;

// sum over the 32 row lanes of a column (threadIdx.y), result broadcast to all of them
__device__ __forceinline__ float column_sum(float v, float (*s)[kCols + 1]) {
  s[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int j = 0; j < kLanes; ++j) t += s[j][threadIdx.x];
  __syncthreads();
  return t;
}

__device__ __forceinline__ float total_count(const float* __restrict__ counts, int R, float (*s)[kCols + 1]) {
  float c = 0.f;
  for (int r = threadIdx.y * kCols + threadIdx.x; r < R; r += kCols * kLanes) c += counts[r];
  // all 1024 threads hold partials: reduce columns first, then across the 32 column sums (same value for every thread)
  const float col = column_sum(c, s);
  s[0][threadIdx.x] = col;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int j = 0; j < kCols; ++j) t += s[0][j];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kCols * kLanes) k_wbn_fwd(
    const float* __restrict__ x, const float* __restrict__ counts, const float* __restrict__ gamma,
    const float* __restrict__ beta, int R, int F, float eps, float momentum, int training, float* running_mean,
    float* running_var, int64_t* tracked, float* __restrict__ y, float* __restrict__ save_mean,
    float* __restrict__ save_rstd) {
  __shared__ float s[kLanes][kCols + 1];
  const int f = blockIdx.x * kCols + threadIdx.x;
  const bool ok = f < F;
  float mean, var, n = 0.f;
  if (training) {
    n = total_count(counts, R, s);
    float a = 0.f;
    if (ok)
      for (int r = threadIdx.y; r < R; r += kLanes) a = fmaf(counts[r], x[(int64_t)r * F + f], a);
    mean = column_sum(a, s) / n;
    float b = 0.f;
    if (ok)
      for (int r = threadIdx.y; r < R; r += kLanes) {
        const float d = x[(int64_t)r * F + f] - mean;
        b = fmaf(counts[r] * d, d, b);
      }
    var = column_sum(b, s) / n;
  } else {
    mean = ok ? running_mean[f] : 0.f;
    var = ok ? running_var[f] : 1.f;
  }
  const float rstd = rsqrtf(var + eps);
  if (ok) {
    const float g = gamma ? gamma[f] * rstd : rstd, b0 = beta ? beta[f] : 0.f;
    for (int r = threadIdx.y; r < R; r += kLanes) y[(int64_t)r * F + f] = fmaf(x[(int64_t)r * F + f] - mean, g, b0);
    if (threadIdx.y == 0) {
      save_mean[f] = mean;
      save_rstd[f] = rstd;
      if (training && running_mean) {
        running_mean[f] = (1.f - momentum) * running_mean[f] + momentum * mean;
        running_var[f] = (1.f - momentum) * running_var[f] + momentum * var * (n / (n - 1.f));
      }
    }
  }
  if (training && tracked && blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0) *tracked += 1;
}

__global__ void __launch_bounds__(kCols * kLanes) k_wbn_bwd(
    const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ counts,
    const float* __restrict__ gamma, const float* __restrict__ save_mean, const float* __restrict__ save_rstd, int R,
    int F, int training, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s[kLanes][kCols + 1];
  const int f = blockIdx.x * kCols + threadIdx.x;
  const bool ok = f < F;
  const float n = training ? total_count(counts, R, s) : 1.f;
  const float mean = ok ? save_mean[f] : 0.f, rstd = ok ? save_rstd[f] : 0.f;
  float a = 0.f, b = 0.f;
  if (ok)
    for (int r = threadIdx.y; r < R; r += kLanes) {
      const float g = dy[(int64_t)r * F + f];
      a += g;
      b = fmaf(g, (x[(int64_t)r * F + f] - mean) * rstd, b);
    }
  const float db = column_sum(a, s), dg = column_sum(b, s);
  if (!ok) return;
  const float sc = rstd * (gamma ? gamma[f] : 1.f);
  for (int r = threadIdx.y; r < R; r += kLanes) {
    const float g = dy[(int64_t)r * F + f];
    float v = g;
    if (training) {
      const float w = counts[r] / n, xh = (x[(int64_t)r * F + f] - mean) * rstd;
      v = g - w * db - w * xh * dg;
    }
    dx[(int64_t)r * F + f] = sc * v;
  }
  if (threadIdx.y == 0) {
    if (dgamma) dgamma[f] = dg;
    if (dbeta) dbeta[f] = db;
  }
}


// ---------------------------------------------------------------------------------------------
// Species-pair coefficient table before its BatchNorm (embed.py:234-249 on the (Zd x Zd) pair table):
//   pre[s, t, o, k] = fe[t, o, k] * (1 + za[s, k] + zb[t, k])
// with fe (Zd, O, K) the per-orbital electron embedding of the TARGET element after f_e, za / zb (Zd, K) the two
// halves of f_z applied to the source / target element embedding.  One kernel forward, one backward (three fixed-order
// sums over the 5.6 MB tensor, L2 resident) instead of ~25 broadcast / reduce launches of the eager expression.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_outer_fwd(const float* __restrict__ fe, const float* __restrict__ za,
                                                        const float* __restrict__ zb, int Zd, int O, int K, float* __restrict__ pre) {
  const int64_t n4 = (int64_t)Zd * Zd * O * (K / 4);
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const int k = (int)(i % (K / 4)) * 4;
    const int64_t r = i / (K / 4);
    const int o = (int)(r % O);
    const int t = (int)((r / O) % Zd), s = (int)(r / ((int64_t)O * Zd));
    const float4 f = ldg4(fe + ((int64_t)t * O + o) * K + k), a = ldg4(za + (int64_t)s * K + k), b = ldg4(zb + (int64_t)t * K + k);
    st4(pre + i * 4, make_float4(f.x * (1.f + a.x + b.x), f.y * (1.f + a.y + b.y), f.z * (1.f + a.z + b.z), f.w * (1.f + a.w + b.w)));
  }
}

// thread = one output element (k fastest: coalesced), sequential fixed-order sums: deterministic.
//   d_fe[t,o,k] = sum_s dpre[s,t,o,k] (1 + za[s,k] + zb[t,k]);  d_za[s,k] = sum_{t,o} dpre fe[t,o,k];  d_zb[t,k] = sum_{s,o} dpre fe[t,o,k]
__global__ void __launch_bounds__(128) k_pair_outer_bwd(const float* __restrict__ dpre, const float* __restrict__ fe,
                                                        const float* __restrict__ za, const float* __restrict__ zb, int Zd, int O,
                                                        int K, float* __restrict__ d_fe, float* __restrict__ d_za,
                                                        float* __restrict__ d_zb) {
  const int64_t n_fe = (int64_t)Zd * O * K, n_z = (int64_t)Zd * K;
  const int64_t i = blockIdx.x * 128ll + threadIdx.x;
  const int64_t OK = (int64_t)O * K, row = (int64_t)Zd * OK;  // strides of t and s in dpre
  if (i < n_fe) {
    const int k = (int)(i % K), t = (int)(i / OK);
    const float b1 = 1.f + zb[(int64_t)t * K + k];
    float acc = 0.f;
#pragma unroll 8
    for (int s = 0; s < Zd; ++s) acc = fmaf(dpre[s * row + i], b1 + za[(int64_t)s * K + k], acc);
    d_fe[i] = acc;
  } else if (i < n_fe + n_z) {
    const int64_t j = i - n_fe;
    const int k = (int)(j % K), s = (int)(j / K);
    float acc = 0.f;
#pragma unroll 8
    for (int64_t to = 0; to < (int64_t)Zd * O; ++to) acc = fmaf(dpre[s * row + to * K + k], fe[to * K + k], acc);
    d_za[j] = acc;
  } else if (i < n_fe + 2 * n_z) {
    const int64_t j = i - n_fe - n_z;
    const int k = (int)(j % K), t = (int)(j / K);
    float acc = 0.f;
#pragma unroll 4
    for (int s = 0; s < Zd; ++s)
      for (int o = 0; o < O; ++o) acc = fmaf(dpre[s * row + ((int64_t)t * O + o) * K + k], fe[((int64_t)t * O + o) * K + k], acc);
    d_zb[j] = acc;
  }
}

}  // namespace

extern "C" int lcao_table_norm_fwd(const float* x, const float* counts, const float* gamma, const float* beta, int64_t R,
                                   int32_t F, float eps, float momentum, int32_t training, float* running_mean,
                                   float* running_var, int64_t* num_batches_tracked, float* y, float* save_mean,
                                   float* save_rstd, void* stream) {
  if (R == 0 || F == 0) return LCAO_OK;
  LCAO_REQUIRE(x && y && save_mean && save_rstd && R < (1ll << 31), "lcao_table_norm_fwd: null buffer");
  LCAO_REQUIRE(training ? counts != nullptr : (running_mean && running_var),
               "lcao_table_norm_fwd: training needs counts, evaluation needs the running statistics");
  LCAO_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "lcao_table_norm_fwd: pass both running buffers or neither");
  k_wbn_fwd<<<(unsigned)((F + kCols - 1) / kCols), dim3(kCols, kLanes), 0, (cudaStream_t)stream>>>(
      x, counts, gamma, beta, (int)R, F, eps, momentum, training, running_mean, running_var, num_batches_tracked, y, save_mean,
      save_rstd);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_table_norm_bwd(const float* dy, const float* x, const float* counts, const float* gamma,
                                   const float* save_mean, const float* save_rstd, int64_t R, int32_t F, int32_t training,
                                   float* dx, float* dgamma, float* dbeta, void* stream) {
  if (R == 0 || F == 0) return LCAO_OK;
  LCAO_REQUIRE(dy && x && save_mean && save_rstd && dx && (!training || counts) && R < (1ll << 31),
               "lcao_table_norm_bwd: null buffer");
  k_wbn_bwd<<<(unsigned)((F + kCols - 1) / kCols), dim3(kCols, kLanes), 0, (cudaStream_t)stream>>>(
      dy, x, counts, gamma, save_mean, save_rstd, (int)R, F, training, dx, dgamma, dbeta);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_pair_outer_fwd(const float* fe, const float* za, const float* zb, int32_t Zd, int32_t O, int32_t K, float* pre,
                                   void* stream) {
  LCAO_REQUIRE(fe && za && zb && pre && Zd > 0 && O > 0 && K > 0 && K % 4 == 0, "lcao_pair_outer_fwd: need K %% 4 == 0 and non-null buffers");
  LCAO_REQUIRE(((reinterpret_cast<uintptr_t>(fe) | reinterpret_cast<uintptr_t>(za) | reinterpret_cast<uintptr_t>(zb) |
                 reinterpret_cast<uintptr_t>(pre)) & 15u) == 0, "lcao_pair_outer_fwd: buffers must be 16-byte aligned");
  const int64_t n4 = (int64_t)Zd * Zd * O * (K / 4);
  const int64_t want = ceil_div64(n4, 256);
  k_pair_outer_fwd<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, (cudaStream_t)stream>>>(fe, za, zb, Zd, O, K, pre);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_pair_outer_bwd(const float* dpre, const float* fe, const float* za, const float* zb, int32_t Zd, int32_t O,
                                   int32_t K, float* d_fe, float* d_za, float* d_zb, void* stream) {
  LCAO_REQUIRE(dpre && fe && za && zb && d_fe && d_za && d_zb && Zd > 0 && O > 0 && K > 0, "lcao_pair_outer_bwd: null buffer");
  const int64_t n = (int64_t)Zd * O * K + 2ll * Zd * K;
  k_pair_outer_bwd<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(dpre, fe, za, zb, Zd, O, K, d_fe, d_za, d_zb);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}
