// C-ABI dispatch for the dense layers (nn/base.py:11-81; call sites SURVEY.md §8 a-7).
#include "common.cuh"

int lcao_simt_linear_fwd(const float*, int64_t, const float*, const float*, float*, int64_t, float*, int64_t, int64_t,
                         int32_t, int32_t, int32_t, cudaStream_t);
int lcao_simt_linear_dgrad(const float*, int64_t, const float*, float*, int64_t, int64_t, int32_t, int32_t, int32_t,
                           cudaStream_t);
int lcao_simt_linear_wgrad(const float*, int64_t, const float*, int64_t, float*, float*, int64_t, int32_t, int32_t,
                           cudaStream_t);

extern "C" int lcao_linear_fwd(const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy,
                               float* pre, int64_t ldp, int64_t M, int32_t K, int32_t Nout, int32_t act, int32_t mode,
                               void* stream) {
  if (M == 0 || Nout == 0) return LCAO_OK;
  LCAO_REQUIRE(X && W && Y && K > 0, "lcao_linear_fwd: null buffer");
  LCAO_REQUIRE(act == LCAO_ACT_NONE || act == LCAO_ACT_SILU, "lcao_linear_fwd: unsupported activation %d", act);
  (void)mode;
  return lcao_simt_linear_fwd(X, ldx, W, bias, Y, ldy, pre, ldp, M, K, Nout, act, (cudaStream_t)stream);
}

extern "C" int lcao_linear_dgrad(const float* dY, int64_t ldy, const float* W, float* dX, int64_t ldx, int64_t M,
                                 int32_t K, int32_t Nout, int32_t accumulate, int32_t mode, void* stream) {
  if (M == 0 || K == 0) return LCAO_OK;
  LCAO_REQUIRE(dY && W && dX && Nout > 0, "lcao_linear_dgrad: null buffer");
  (void)mode;
  return lcao_simt_linear_dgrad(dY, ldy, W, dX, ldx, M, K, Nout, accumulate, (cudaStream_t)stream);
}

extern "C" int lcao_linear_wgrad(const float* dY, int64_t ldy, const float* X, int64_t ldx, float* dW, float* db,
                                 int64_t M, int32_t K, int32_t Nout, int32_t mode, void* stream) {
  if (M == 0 || K == 0 || Nout == 0) return LCAO_OK;
  LCAO_REQUIRE(dY && X && dW, "lcao_linear_wgrad: null buffer");
  (void)mode;
  return lcao_simt_linear_wgrad(dY, ldy, X, ldx, dW, db, M, K, Nout, (cudaStream_t)stream);
}
