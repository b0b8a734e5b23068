// FP32 CUDA-core GEMM for the dense layers (exact-fp32 mode and odd shapes).  The tcgen05 path in
// gemm_tc.cu takes over for the large edge-sized layers; this kernel is the numerically exact
// companion (fp32 FMA, fp32 accumulate) that parity tests and tiny / unaligned layers use.
//   C[m,n] = sum_k A(m,k) * B(n,k)
// forward : A = X (M,K) k-contiguous, B = W (Nout,K) k-contiguous, fused bias + SiLU epilogue
// dgrad   : A = dY (M,Nout) k-contiguous, B = W read as (k'=n, n'=k) n-contiguous
// wgrad   : A = dY read as (m'=n, k'=m) m-contiguous, B = X read as (n'=k, k'=m) n-contiguous,
//           split over the (huge) row dimension with atomic accumulation into dW.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, GT = 256;

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  float* pre; int64_t ldp;
  const float* bias;
  int64_t M; int N; int64_t K;   // C is M x N, contraction length K
  int64_t k_chunk;               // contraction range per blockIdx.z
  int act, accumulate, atomic_out, vecA, vecB, vecC;
};

// stage a (rows x BK) tile of a k-contiguous operand, transposed into sm[k][row]
__device__ __forceinline__ void load_kc(const float* __restrict__ P, int64_t ld, int64_t row0, int64_t nrows, int64_t k0,
                                        int64_t kend, bool vec, float (*sm)[BM + PAD]) {
  const int t = threadIdx.x;
  const int r = t >> 1, kq = (t & 1) * 8;
  const int64_t row = row0 + r;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t k = k0 + kq + h * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows) {
      if (vec && k + 3 < kend) {
        v = ldg4(P + row * ld + k);
      } else {
        if (k < kend) v.x = __ldg(P + row * ld + k);
        if (k + 1 < kend) v.y = __ldg(P + row * ld + k + 1);
        if (k + 2 < kend) v.z = __ldg(P + row * ld + k + 2);
        if (k + 3 < kend) v.w = __ldg(P + row * ld + k + 3);
      }
    }
    sm[kq + h * 4 + 0][r] = v.x; sm[kq + h * 4 + 1][r] = v.y; sm[kq + h * 4 + 2][r] = v.z; sm[kq + h * 4 + 3][r] = v.w;
  }
}

// stage a (BK x cols) tile of an operand stored (k, col) col-contiguous
__device__ __forceinline__ void load_nc(const float* __restrict__ P, int64_t ld, int64_t col0, int64_t ncols, int64_t k0,
                                        int64_t kend, bool vec, float (*sm)[BM + PAD]) {
  const int t = threadIdx.x;
  const int kk = t >> 4, cq = (t & 15) * 8;
  const int64_t k = k0 + kk;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t col = col0 + cq + h * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < kend) {
      if (vec && col + 3 < ncols) {
        v = ldg4(P + k * ld + col);
      } else {
        if (col < ncols) v.x = __ldg(P + k * ld + col);
        if (col + 1 < ncols) v.y = __ldg(P + k * ld + col + 1);
        if (col + 2 < ncols) v.z = __ldg(P + k * ld + col + 2);
        if (col + 3 < ncols) v.w = __ldg(P + k * ld + col + 3);
      }
    }
    *reinterpret_cast<float4*>(&sm[kk][cq + h * 4]) = v;
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GT) k_gemm(const GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int64_t n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  const int64_t kend = min(g.K, kbeg + g.k_chunk);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    if (A_KC) load_kc(g.A, g.lda, m0, g.M, k0, kend, g.vecA, As);
    else load_nc(g.A, g.lda, m0, g.M, k0, kend, g.vecA, As);
    if (B_KC) load_kc(g.B, g.ldb, n0, g.N, k0, kend, g.vecB, Bs);
    else load_nc(g.B, g.ldb, n0, g.N, k0, kend, g.vecB, Bs);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int64_t n = n0 + (jh ? 64 : 0) + tx * 4;
      if (n >= g.N) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (g.atomic_out) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < g.N) atomicAdd(g.C + m * g.ldc + n + j, v[j]);
        continue;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j >= g.N) continue;
        if (g.bias) v[j] += __ldg(g.bias + n + j);
        if (g.accumulate) v[j] += g.C[m * g.ldc + n + j];
      }
      if (g.vecC && n + 3 < g.N) {
        if (g.pre) st4(g.pre + m * g.ldp + n, make_float4(v[0], v[1], v[2], v[3]));
        if (g.act != LCAO_ACT_NONE) { v[0] = act_fwdf(g.act, v[0]); v[1] = act_fwdf(g.act, v[1]); v[2] = act_fwdf(g.act, v[2]); v[3] = act_fwdf(g.act, v[3]); }
        st4(g.C + m * g.ldc + n, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j >= g.N) continue;
          if (g.pre) g.pre[m * g.ldp + n + j] = v[j];
          g.C[m * g.ldc + n + j] = act_fwdf(g.act, v[j]);
        }
      }
    }
  }
}

// db[n] += sum_m dY[m,n] over a 512-row slab per CTA
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ dY, int64_t ldy, int64_t M, int N,
                                                float* __restrict__ db) {
  const int64_t r0 = (int64_t)blockIdx.x * 512, r1 = min(M, r0 + 512);
  for (int n = threadIdx.x; n < N; n += 256) {
    float s = 0.f;
    for (int64_t m = r0; m < r1; ++m) s += __ldg(dY + m * ldy + n);
    atomicAdd(db + n, s);
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// Tiny row counts (species / pair tables: M = 37 ... 512).  The tiled kernel above runs them as one or two
// CTAs stepping through K with a barrier per 16 columns (~30 us of pure latency); here the OUTPUT columns are
// spread over the grid (8 per CTA, their operand columns staged in shared memory), one thread per row, all
// loads of a row in flight at once.  FP32 FMA, same summation order over k as a plain dot product.
// ---------------------------------------------------------------------------------------------
constexpr int kTinyCols = 8, kTinyMaxK = 256;

// Y[m, n0..n0+7] = act(sum_k X[m,k] * Wt(n,k) + bias) ; TRANS: Wt(n,k) = W[k*ldw + n] (dgrad) else W[n*ldw + k]
template <bool TRANS>
__global__ void __launch_bounds__(128) k_tiny_rows(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W,
                                                   int64_t ldw, const float* __restrict__ bias, float* __restrict__ Y,
                                                   int64_t ldy, float* __restrict__ pre, int64_t ldp, int M, int K, int N,
                                                   int act, int accumulate) {
  __shared__ float s_w[kTinyCols][kTinyMaxK];
  const int n0 = blockIdx.x * kTinyCols;
  for (int i = threadIdx.x; i < kTinyCols * K; i += 128) {
    const int j = TRANS ? i % kTinyCols : i / K, k = TRANS ? i / kTinyCols : i % K;
    float w = 0.f;
    if (n0 + j < N) w = TRANS ? __ldg(W + (int64_t)k * ldw + n0 + j) : __ldg(W + (int64_t)(n0 + j) * ldw + k);
    s_w[j][k] = w;
  }
  __syncthreads();
  for (int m = blockIdx.y * 128 + threadIdx.x; m < M; m += gridDim.y * 128) {
    float acc[kTinyCols];
#pragma unroll
    for (int j = 0; j < kTinyCols; ++j) acc[j] = 0.f;
    const float* xr = X + (int64_t)m * ldx;
    for (int k = 0; k < K; ++k) {
      const float x = __ldg(xr + k);
#pragma unroll
      for (int j = 0; j < kTinyCols; ++j) acc[j] = fmaf(x, s_w[j][k], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kTinyCols; ++j) {
      const int n = n0 + j;
      if (n >= N) continue;
      float v = acc[j];
      if (bias) v += __ldg(bias + n);
      if (accumulate) v += Y[(int64_t)m * ldy + n];
      if (pre) pre[(int64_t)m * ldp + n] = v;
      Y[(int64_t)m * ldy + n] = act_fwdf(act, v);
    }
  }
}

// dW[n, k] += sum_m dY[m,n] X[m,k] ; db[n] += sum_m dY[m,n].  CTA = one output row n, thread = column k.
__global__ void __launch_bounds__(128) k_tiny_wgrad(const float* __restrict__ dY, int64_t ldy, const float* __restrict__ X,
                                                    int64_t ldx, float* __restrict__ dW, float* __restrict__ db, int M, int K) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += 128) {
    float acc = 0.f;
    for (int m = 0; m < M; ++m) acc = fmaf(__ldg(dY + (int64_t)m * ldy + n), __ldg(X + (int64_t)m * ldx + k), acc);
    dW[(int64_t)n * K + k] += acc;
  }
  if (db && threadIdx.x == 0) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += __ldg(dY + (int64_t)m * ldy + n);
    db[n] += s;
  }
}

inline bool tiny_ok(int64_t M, int K) { return M <= 512 && K <= kTinyMaxK; }

}  // namespace

int lcao_simt_linear_fwd(const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy,
                         float* pre, int64_t ldp, int64_t M, int32_t K, int32_t Nout, int32_t act, cudaStream_t st) {
  if (tiny_ok(M, K)) {
    dim3 grid((unsigned)ceil_div64(Nout, kTinyCols), (unsigned)ceil_div64(M, 128));
    k_tiny_rows<false><<<grid, 128, 0, st>>>(X, ldx, W, K, bias, Y, ldy, pre, ldp, (int)M, K, Nout, act, 0);
    LCAO_LAUNCH_CHECK();
    return LCAO_OK;
  }
  GemmArgs g{};
  g.A = X; g.lda = ldx; g.B = W; g.ldb = K; g.C = Y; g.ldc = ldy; g.pre = pre; g.ldp = ldp; g.bias = bias;
  g.M = M; g.N = Nout; g.K = K; g.k_chunk = K; g.act = act;
  g.vecA = al16(X) && ldx % 4 == 0;
  g.vecB = al16(W) && K % 4 == 0;
  g.vecC = al16(Y) && ldy % 4 == 0 && (!pre || (al16(pre) && ldp % 4 == 0));
  dim3 grid((unsigned)ceil_div64(Nout, BN), (unsigned)ceil_div64(M, BM), 1);
  k_gemm<true, true><<<grid, GT, 0, st>>>(g);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

int lcao_simt_linear_dgrad(const float* dY, int64_t ldy, const float* W, float* dX, int64_t ldx, int64_t M, int32_t K,
                           int32_t Nout, int32_t accumulate, cudaStream_t st) {
  if (tiny_ok(M, Nout)) {  // dX[m, k] = sum_n dY[m,n] W[n,k]: contraction over Nout, "weight" read transposed
    dim3 grid((unsigned)ceil_div64(K, kTinyCols), (unsigned)ceil_div64(M, 128));
    k_tiny_rows<true><<<grid, 128, 0, st>>>(dY, ldy, W, K, nullptr, dX, ldx, nullptr, 0, (int)M, Nout, K, LCAO_ACT_NONE, accumulate);
    LCAO_LAUNCH_CHECK();
    return LCAO_OK;
  }
  GemmArgs g{};
  g.A = dY; g.lda = ldy; g.B = W; g.ldb = K; g.C = dX; g.ldc = ldx;
  g.M = M; g.N = K; g.K = Nout; g.k_chunk = Nout; g.accumulate = accumulate;
  g.vecA = al16(dY) && ldy % 4 == 0;
  g.vecB = al16(W) && K % 4 == 0;
  g.vecC = al16(dX) && ldx % 4 == 0;
  dim3 grid((unsigned)ceil_div64(K, BN), (unsigned)ceil_div64(M, BM), 1);
  k_gemm<true, false><<<grid, GT, 0, st>>>(g);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

int lcao_simt_linear_wgrad(const float* dY, int64_t ldy, const float* X, int64_t ldx, float* dW, float* db, int64_t M,
                           int32_t K, int32_t Nout, cudaStream_t st) {
  if (M <= 64) {  // (longer row loops are better served by the split-K tiled kernel below)
    k_tiny_wgrad<<<(unsigned)Nout, 128, 0, st>>>(dY, ldy, X, ldx, dW, db, (int)M, K);
    LCAO_LAUNCH_CHECK();
    return LCAO_OK;
  }
  GemmArgs g{};
  g.A = dY; g.lda = ldy; g.B = X; g.ldb = ldx; g.C = dW; g.ldc = K;
  g.M = Nout; g.N = K; g.K = M; g.atomic_out = 1;
  g.vecA = al16(dY) && ldy % 4 == 0;
  g.vecB = al16(X) && ldx % 4 == 0;
  const int64_t tiles = ceil_div64(Nout, BM) * ceil_div64(K, BN);
  int64_t splits = ceil_div64(2 * 148, tiles);
  const int64_t max_splits = ceil_div64(M, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  g.k_chunk = ceil_div64(ceil_div64(M, splits), BK) * BK;
  splits = ceil_div64(M, g.k_chunk);
  dim3 grid((unsigned)ceil_div64(K, BN), (unsigned)ceil_div64(Nout, BM), (unsigned)splits);
  k_gemm<false, false><<<grid, GT, 0, st>>>(g);
  LCAO_LAUNCH_CHECK();
  if (db) {
    k_colsum<<<(unsigned)ceil_div64(M, 512), 256, 0, st>>>(dY, ldy, M, Nout, db);
    LCAO_LAUNCH_CHECK();
  }
  return LCAO_OK;
}
