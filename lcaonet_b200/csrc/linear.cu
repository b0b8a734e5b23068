// C-ABI dispatch for the dense layers (nn/base.py:11-81; call sites SURVEY.md §8 a-7):
// tcgen05 3xTF32 / TF32 kernels (gemm_tc.cu) for the large row-streaming layers, FP32 CUDA-core
// kernels (gemm_simt.cu) for exact-fp32 mode and for small or unaligned shapes.
#include "common.cuh"

int lcao_simt_linear_fwd(const float*, int64_t, const float*, const float*, float*, int64_t, float*, int64_t, int64_t,
                         int32_t, int32_t, int32_t, cudaStream_t);
int lcao_simt_linear_dgrad(const float*, int64_t, const float*, float*, int64_t, int64_t, int32_t, int32_t, int32_t,
                           cudaStream_t);
int lcao_simt_linear_wgrad(const float*, int64_t, const float*, int64_t, float*, float*, int64_t, int32_t, int32_t,
                           cudaStream_t);
bool lcao_tc_rows_ok(int64_t M, int Kc, int Nb, int64_t lda, int64_t ldy, const void* A, const void* Y);
int lcao_tc_rows(const float* A, int64_t lda, const float* W, int64_t ldw, int b_trans, const float* bias, const float* G,
                 int64_t ldg, float* Y, int64_t ldy, float* pre, int64_t ldp, int64_t M, int Kc, int Nb, int act,
                 int accumulate, int x3, cudaStream_t st);
bool lcao_tc_wgrad_ok(int64_t M, int Kx, int64_t ldy, int64_t ldx, const void* dY, const void* X);
int lcao_tc_wgrad_reduce_batch(const int64_t* desc, int n, cudaStream_t st);
int lcao_tc_wgrad(const float* dY, int64_t ldy, const float* X, int64_t ldx, float* dW, int64_t ldw, float* db, int64_t M,
                  int Kx, int x3, float* part, cudaStream_t st, int64_t* defer_desc);
int64_t lcao_tc_wgrad_scratch(int64_t M, int Kx);

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int imin(int a, int b) { return a < b ? a : b; }

extern "C" int lcao_linear_fwd(const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy,
                               float* pre, int64_t ldp, int64_t M, int32_t K, int32_t Nout, int32_t act, int32_t mode,
                               void* stream) {
  if (M == 0 || Nout == 0) return LCAO_OK;
  LCAO_REQUIRE(X && W && Y && K > 0, "lcao_linear_fwd: null buffer");
  LCAO_REQUIRE(act >= LCAO_ACT_NONE && act <= LCAO_ACT_LAST, "lcao_linear_fwd: unsupported activation %d", act);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = mode != LCAO_GEMM_FP32 && Nout % 16 == 0 && K <= 128 && al16(W) && (!bias || al16(bias)) &&
                  (!pre || (al16(pre) && ldp % 4 == 0)) && lcao_tc_rows_ok(M, K, imin(Nout, 128), ldx, ldy, X, Y);
  if (!tc) return lcao_simt_linear_fwd(X, ldx, W, bias, Y, ldy, pre, ldp, M, K, Nout, act, st);
  // the tcgen05 epilogue fuses SiLU only (the reference default); the other activations run as one elementwise pass
  // over the GEMM's output, which keeps the hot kernel free of their code
  const bool post = act != LCAO_ACT_NONE && act != LCAO_ACT_SILU;
  float* out = (post && pre) ? pre : Y;
  const int64_t ldo = (post && pre) ? ldp : ldy;
  for (int n0 = 0; n0 < Nout; n0 += 128) {
    const int nb = imin(128, Nout - n0);
    int rc = lcao_tc_rows(X, ldx, W + (int64_t)n0 * K, K, 0, bias ? bias + n0 : nullptr, nullptr, 0, out + n0, ldo,
                          (pre && !post) ? pre + n0 : nullptr, ldp, M, K, nb, post ? LCAO_ACT_NONE : act, 0,
                          mode == LCAO_GEMM_TF32X3, st);
    if (rc) return rc;
  }
  if (post) return lcao_act_fwd(out, ldo, Y, ldy, M, Nout, act, stream);
  return LCAO_OK;
}

// dY * act'(H) is materialised once into the caller's scratch (all GEMM kernels take it as a plain operand)
static int act_bwd_to_scratch(const float*& dY, int64_t& ldy, const float* H, int64_t ldh, int32_t act, int64_t M,
                              int32_t Nout, float* scratch, void* stream, const char* who) {
  if (act == LCAO_ACT_NONE || !H) return LCAO_OK;
  LCAO_REQUIRE(scratch, "%s: an (M, Nout) scratch buffer is needed for dY * act'(H)", who);
  int rc = lcao_act_bwd(dY, ldy, H, ldh, scratch, Nout, M, Nout, act, stream);
  if (rc) return rc;
  dY = scratch;
  ldy = Nout;
  return LCAO_OK;
}

static bool dgrad_tc(const float* dY, int64_t ldy, const float* W, const float* dX, int64_t ldx, int64_t M, int32_t K,
                     int32_t Nout, int32_t mode) {
  return mode != LCAO_GEMM_FP32 && Nout % 32 == 0 && K % 16 == 0 && al16(W) &&
         lcao_tc_rows_ok(M, imin(Nout, 128), imin(K, 128), ldy, ldx, dY, dX);
}
static bool wgrad_tc(const float* dY, int64_t ldy, const float* X, int64_t ldx, int64_t M, int32_t K, int32_t Nout,
                     int32_t mode) {
  return mode != LCAO_GEMM_FP32 && Nout % 128 == 0 && lcao_tc_wgrad_ok(M, K, ldy, ldx, dY, X);
}

extern "C" int64_t lcao_linear_bwd_scratch(const float* dY, int64_t ldy, const float* H, int64_t ldh, int32_t act,
                                           const float* W, const float* X, int64_t ldx, const float* dX, int64_t lddx,
                                           int64_t M, int32_t K, int32_t Nout, int32_t mode) {
  (void)ldh; (void)W; (void)dX; (void)lddx;
  // every GEMM kernel takes dY * act'(H) as a plain operand: one elementwise pass into scratch; the tcgen05
  // weight-gradient kernel additionally needs room for its per-CTA partial tiles (deterministic reduction)
  int64_t n = (act == LCAO_ACT_NONE || !H) ? 0 : M * (int64_t)Nout;
  if (X && wgrad_tc(dY, ldy, X, ldx, M, K, Nout, mode)) n += lcao_tc_wgrad_scratch(M, K);
  return n;
}

extern "C" int lcao_linear_dgrad(const float* dY, int64_t ldy, const float* H, int64_t ldh, int32_t act, const float* W,
                                 float* dX, int64_t ldx, int64_t M, int32_t K, int32_t Nout, int32_t accumulate,
                                 int32_t mode, float* scratch, void* stream) {
  if (M == 0 || K == 0) return LCAO_OK;
  LCAO_REQUIRE(dY && W && dX && Nout > 0, "lcao_linear_dgrad: null buffer");
  LCAO_REQUIRE(act >= LCAO_ACT_NONE && act <= LCAO_ACT_LAST, "lcao_linear_dgrad: unsupported activation %d", act);
  cudaStream_t st = (cudaStream_t)stream;
  {
    int rc = act_bwd_to_scratch(dY, ldy, H, ldh, act, M, Nout, scratch, stream, "lcao_linear_dgrad");
    if (rc) return rc;
  }
  if (!dgrad_tc(dY, ldy, W, dX, ldx, M, K, Nout, mode))
    return lcao_simt_linear_dgrad(dY, ldy, W, dX, ldx, M, K, Nout, accumulate, st);
  for (int n0 = 0; n0 < K; n0 += 128) {          // output columns
    const int nb = imin(128, K - n0);
    for (int c0 = 0; c0 < Nout; c0 += 128) {     // contraction chunks
      const int kc = imin(128, Nout - c0);
      int rc = lcao_tc_rows(dY + c0, ldy, W + (int64_t)c0 * K + n0, K, 1, nullptr, nullptr, 0, dX + n0, ldx, nullptr, 0, M,
                            kc, nb, LCAO_ACT_NONE, accumulate || c0 > 0, mode == LCAO_GEMM_TF32X3, st);
      if (rc) return rc;
    }
  }
  return LCAO_OK;
}


// dX = (dY W) * act'(G): the data gradient of a layer whose INPUT came out of an activation with pre-activation G.
// The tcgen05 kernel applies the factor in its epilogue (last contraction chunk), which saves the separate
// lcao_act_bwd pass (one read + one write of M x K) that would follow lcao_linear_dgrad.
extern "C" int lcao_linear_dgrad_act(const float* dY, int64_t ldy, const float* W, const float* G, int64_t ldg, int32_t act,
                                     float* dX, int64_t ldx, int64_t M, int32_t K, int32_t Nout, int32_t mode,
                                     void* stream) {
  if (M == 0 || K == 0) return LCAO_OK;
  LCAO_REQUIRE(dY && W && dX && G && Nout > 0, "lcao_linear_dgrad_act: null buffer");
  LCAO_REQUIRE(act > LCAO_ACT_NONE && act <= LCAO_ACT_LAST, "lcao_linear_dgrad_act: unsupported activation %d", act);
  cudaStream_t st = (cudaStream_t)stream;
  if (!dgrad_tc(dY, ldy, W, dX, ldx, M, K, Nout, mode) || ldg % 4 != 0 || ((uintptr_t)G & 15)) {
    int rc = lcao_simt_linear_dgrad(dY, ldy, W, dX, ldx, M, K, Nout, 0, st);
    if (rc) return rc;
    return lcao_act_bwd(dX, ldx, G, ldg, dX, ldx, M, K, act, stream);  // elementwise, in place
  }
  if (act != LCAO_ACT_SILU) {  // (the tcgen05 epilogue fuses SiLU' only)
    int rc = lcao_linear_dgrad(dY, ldy, nullptr, 0, LCAO_ACT_NONE, W, dX, ldx, M, K, Nout, 0, mode, nullptr, stream);
    if (rc) return rc;
    return lcao_act_bwd(dX, ldx, G, ldg, dX, ldx, M, K, act, stream);
  }
  for (int n0 = 0; n0 < K; n0 += 128) {          // output columns
    const int nb = imin(128, K - n0);
    for (int c0 = 0; c0 < Nout; c0 += 128) {     // contraction chunks; the factor goes on the last one
      const int kc = imin(128, Nout - c0);
      const bool last = c0 + 128 >= Nout;
      int rc = lcao_tc_rows(dY + c0, ldy, W + (int64_t)c0 * K + n0, K, 1, nullptr, last ? G + n0 : nullptr, ldg, dX + n0, ldx,
                            nullptr, 0, M, kc, nb, LCAO_ACT_NONE, c0 > 0, mode == LCAO_GEMM_TF32X3, st);
      if (rc) return rc;
    }
  }
  return LCAO_OK;
}

extern "C" int lcao_linear_wgrad(const float* dY, int64_t ldy, const float* H, int64_t ldh, int32_t act, const float* X,
                                 int64_t ldx, float* dW, float* db, int64_t M, int32_t K, int32_t Nout, int32_t mode,
                                 float* scratch, void* stream) {
  if (M == 0 || K == 0 || Nout == 0) return LCAO_OK;
  LCAO_REQUIRE(dY && X && dW, "lcao_linear_wgrad: null buffer");
  LCAO_REQUIRE(act >= LCAO_ACT_NONE && act <= LCAO_ACT_LAST, "lcao_linear_wgrad: unsupported activation %d", act);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = wgrad_tc(dY, ldy, X, ldx, M, K, Nout, mode);  // decided on the caller's dY, like the scratch query
  float* part = (act == LCAO_ACT_NONE || !H) ? scratch : scratch + M * (int64_t)Nout;
  {
    int rc = act_bwd_to_scratch(dY, ldy, H, ldh, act, M, Nout, scratch, stream, "lcao_linear_wgrad");
    if (rc) return rc;
  }
  if (!tc || !wgrad_tc(dY, ldy, X, ldx, M, K, Nout, mode)) return lcao_simt_linear_wgrad(dY, ldy, X, ldx, dW, db, M, K, Nout, st);
  LCAO_REQUIRE(scratch, "lcao_linear_wgrad: scratch of lcao_linear_bwd_scratch() floats is needed (per-CTA partial tiles)");
  for (int n0 = 0; n0 < Nout; n0 += 128) {
    int rc = lcao_tc_wgrad(dY + n0, ldy, X, ldx, dW + (int64_t)n0 * K, K, db ? db + n0 : nullptr, M, K,
                           mode == LCAO_GEMM_TF32X3, part, st, nullptr);
    if (rc) return rc;
  }
  return LCAO_OK;
}

// The same with the second stage (the sum of the per-CTA partial tiles into dW / db) DEFERRED: a training step needs its
// weight gradients only at the optimizer, so the ~30 small reduction launches of a backward pass can be one
// (lcao_wgrad_reduce_batch).  desc (HOST memory, 6 int64 per 128-column pass, Nout / 128 passes) receives the
// descriptors, *n_desc their number — 0 when the shape took the CUDA-core kernel, which finishes dW at once.  scratch:
// lcao_linear_bwd_scratch() floats PER PASS, and it must stay untouched until lcao_wgrad_reduce_batch has run.
extern "C" int lcao_linear_wgrad_deferred(const float* dY, int64_t ldy, const float* X, int64_t ldx, float* dW, float* db,
                                          int64_t M, int32_t K, int32_t Nout, int32_t mode, float* scratch, int64_t* desc,
                                          int32_t* n_desc, void* stream) {
  LCAO_REQUIRE(desc && n_desc, "lcao_linear_wgrad_deferred: null descriptor buffer");
  *n_desc = 0;
  if (M == 0 || K == 0 || Nout == 0) return LCAO_OK;
  LCAO_REQUIRE(dY && X && dW, "lcao_linear_wgrad_deferred: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (!wgrad_tc(dY, ldy, X, ldx, M, K, Nout, mode)) return lcao_simt_linear_wgrad(dY, ldy, X, ldx, dW, db, M, K, Nout, st);
  LCAO_REQUIRE(scratch, "lcao_linear_wgrad_deferred: scratch is needed (per-CTA partial tiles of every pass)");
  const int64_t per_pass = lcao_tc_wgrad_scratch(M, K);
  for (int n0 = 0; n0 < Nout; n0 += 128) {
    int rc = lcao_tc_wgrad(dY + n0, ldy, X, ldx, dW + (int64_t)n0 * K, K, db ? db + n0 : nullptr, M, K,
                           mode == LCAO_GEMM_TF32X3, scratch + (n0 / 128) * per_pass, st, desc + 6 * (int64_t)*n_desc);
    if (rc) return rc;
    *n_desc += 1;
  }
  return LCAO_OK;
}

extern "C" int lcao_wgrad_reduce_batch(const int64_t* desc, int32_t n, void* stream) {
  if (n <= 0) return LCAO_OK;
  LCAO_REQUIRE(desc, "lcao_wgrad_reduce_batch: null descriptors");
  return lcao_tc_wgrad_reduce_batch(desc, n, (cudaStream_t)stream);
}
