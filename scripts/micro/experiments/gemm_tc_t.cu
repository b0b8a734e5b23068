// Row-streaming tcgen05 GEMM with the WEIGHT resident in tensor memory (sm_100a only).
//
//   Y[M, Nb] = epi( A[M, Kc] * Bop[Kc, Nb] )        forward: Bop(k, n) = W[n*ldw + k] ; dgrad: Bop(k, n) = W[k*ldw + n]
//
// gemm_tc.cu keeps the 3xTF32 weight (hi + lo, 128 KB at K = N = 128) in shared memory, which leaves room for
// only 3 + 3 operand stages: the kernel is bound by the bytes it can keep in flight (profiles/r01_notes.md).
// Here the product is computed TRANSPOSED,
//       Y^T[Nb, rows] = W[Nb, Kc] * A^T[Kc, rows],
// so that the weight is the M-side operand of tcgen05.mma and can live in TMEM (lane = output column n, column =
// k; hi in columns [0,128), lo in [128,256)), written once per CTA with tcgen05.st.  Shared memory then holds
// nothing but the activation ring: 64-row stages of 8 KB (hi) + 8 KB (lo), 12 + 12 deep.  The accumulators take
// the other 256 TMEM columns (2 buffers x {main, correction} x 64 rows).  A side effect of the transposition: an
// epilogue thread owns one OUTPUT COLUMN, so for every row m the 32 lanes of a warp store 32 consecutive floats —
// coalesced 128-byte lines without any staging.
//
// Arithmetic identical to gemm_tc.cu: kind::tf32, FP32 accumulate; 3xTF32 = W_hi*X_hi (main) + W_hi*X_lo + W_lo*X_hi
// (correction accumulator), or W_hi*X_hi only in 1x mode.
// Warp roles (416 threads): 0-3 epilogue (+ weight fill), 4-7 hi/lo split, 8 MMA issuer + TMEM allocator, 9-12 loaders.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kTileRows = 64;                       // activation rows per tile = MMA N
constexpr int kChunk = 32;                          // contraction elements per stage (one 128-byte swizzled row)
constexpr uint32_t kStage = kTileRows * 128;        // 8 KB
constexpr int kThreadsT = (4 + 4 + 1 + 4) * 32;     // 416
constexpr uint32_t kSpin = 1u << 28;
constexpr uint32_t kColLo = 128, kColAcc = 256;     // TMEM column map: W_hi | W_lo | accumulators

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = s32(b);
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && ++spins > kSpin) __trap();  // a protocol bug aborts instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcf_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcf_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa_arrive(uint64_t* b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// same load without the wait: several can be in flight; call tm_wait_ld() before using the values
__device__ __forceinline__ void tm_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ float hi_of(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float sig_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct RowsTArgs {
  const float* A; int64_t lda;
  const float* W; int64_t ldw;
  const float* bias;
  const float* G; int64_t ldg;
  float* Y; int64_t ldy;
  float* pre; int64_t ldp;
  int64_t M; int Kc; int Nb;
  int b_trans, act, accumulate, x3, stages, lo_stages;
  int debug;  // ablation bits (LCAO_TC_DEBUG): 1 no output stores, 2 no input copies, 4 no MMAs, 8 no epilogue work, 16 no split
};

__global__ void __launch_bounds__(kThreadsT, 1) k_tc_rows_t(const RowsTArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = g.stages, L = g.lo_stages;
  uint8_t* sHi = smem;
  uint8_t* sLo = sHi + (size_t)R * kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLo + (size_t)(g.x3 ? L : 0) * kStage);
  uint64_t* raw_full = bars;          // [R] loaders -> split
  uint64_t* full = raw_full + R;      // [R] split -> MMA
  uint64_t* hi_empty = full + R;      // [R] MMA -> loaders
  uint64_t* lo_empty = hi_empty + R;  // [L] MMA -> split
  uint64_t* tfull = lo_empty + L;     // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;       // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int64_t ntiles = (g.M + kTileRows - 1) / kTileRows;
  const int nchunk = g.Kc / kChunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) { mb_init(&raw_full[s], 4 * 32); mb_init(&full[s], 4 * 32); mb_init(&hi_empty[s], 1); }
    for (int s = 0; s < L; ++s) mb_init(&lo_empty[s], 1);
    for (int a = 0; a < 2; ++a) { mb_init(&tfull[a], 1); mb_init(&tempty[a], 4 * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcf_before();
  __syncthreads();
  tcf_after();
  const uint32_t tm = *tmem_slot;

  if (warp >= 9) {
    // ============================== loaders: 64-row stages, 8 lanes per 128-byte row ==============================
    const int lw = warp - 9, row_in = lane >> 3, c = lane & 7;
    const uint32_t base = s32(sHi);
    uint32_t it = 0;
    for (int64_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
      const int64_t mrow = tb * kTileRows + 16 * lw + row_in;
      for (int kc = 0; kc < nchunk; ++kc, ++it) {
        const int s = it % R;
        mb_wait(&hi_empty[s], ((it / R) & 1) ^ 1);
        const uint32_t dst = base + s * kStage;
        const float* src = g.A + mrow * g.lda + kc * kChunk + c * 4;
        if (!(g.debug & 2)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = mrow + 4 * j < g.M;
            cpa16(dst + sw128(16 * lw + 4 * j + row_in, c), ok ? src + (int64_t)4 * j * g.lda : g.A, ok ? 16u : 0u);
          }
        }
        cpa_arrive(&raw_full[s]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================== split: hi in place, lo into its ring ==============================
    const int t = threadIdx.x - 128;
    uint32_t it = 0;
    for (int64_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
      for (int kc = 0; kc < nchunk; ++kc, ++it) {
        const int s = it % R, l = it % L;
        mb_wait(&raw_full[s], (it / R) & 1);
        if (g.x3) mb_wait(&lo_empty[l], ((it / L) & 1) ^ 1);
        if (g.x3 && !(g.debug & 16)) {
          uint8_t* ph = sHi + (size_t)s * kStage + t * 16;
          uint8_t* pl = sLo + (size_t)l * kStage + t * 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(ph + i * 2048);
            const float4 hi = make_float4(hi_of(x.x), hi_of(x.y), hi_of(x.z), hi_of(x.w));
            *reinterpret_cast<float4*>(ph + i * 2048) = hi;
            *reinterpret_cast<float4*>(pl + i * 2048) =
                make_float4(hi_of(x.x - hi.x), hi_of(x.y - hi.y), hi_of(x.z - hi.z), hi_of(x.w - hi.w));
          }
        }
        fence_async();
        mb_arrive(&full[s]);
      }
    }
  } else if (warp == 8) {
    // ============================== MMA issuer ==============================
    asm volatile("bar.sync 1, 160;" ::: "memory");  // the weight is in TMEM (epilogue warps 0-3 + this warp)
    tcf_after();
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, kTileRows);
      const uint32_t hiB = s32(sHi), loB = s32(sLo);
      uint32_t it = 0, tile = 0;
      for (int64_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x, ++tile) {
        const int acc = tile & 1;
        mb_wait(&tempty[acc], ((tile >> 1) & 1) ^ 1);
        tcf_after();
        const uint32_t d = tm + kColAcc + acc * 128, dc = d + 64;
        for (int kc = 0; kc < nchunk; ++kc, ++it) {
          const int s = it % R, l = it % L;
          mb_wait(&full[s], (it / R) & 1);
          tcf_after();
          const uint32_t x_hi = hiB + s * kStage, x_lo = loB + l * kStage;
#pragma unroll
          for (int kk = 0; kk < ((g.debug & 4) ? 0 : kChunk / 8); ++kk) {
            const uint32_t w_hi = tm + kc * kChunk + kk * 8, w_lo = w_hi + kColLo;
            const uint64_t dXh = desc_sw128(x_hi + kk * 32);
            umma_ts(d, w_hi, dXh, idesc, (kc | kk) != 0);
            if (g.x3) {
              umma_ts(dc, w_hi, desc_sw128(x_lo + kk * 32), idesc, (kc | kk) != 0);
              umma_ts(dc, w_lo, dXh, idesc, 1);
            }
          }
          umma_commit(&hi_empty[s]);
          if (g.x3) umma_commit(&lo_empty[l]);
        }
        umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue warps ==============================
    // ---- first: this thread's weight row (output column n) into TMEM, split hi / lo
    const int n = warp * 32 + lane;
    for (int kb = 0; kb < nchunk; ++kb) {
      float w[32], hi[32];
      if (n < g.Nb) {
        if (!g.b_trans) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = ldg4(g.W + (int64_t)n * g.ldw + kb * 32 + q * 4);
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) w[k] = __ldg(g.W + (int64_t)(kb * 32 + k) * g.ldw + n);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) w[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) hi[k] = hi_of(w[k]);
      tm_st32(tm + ((uint32_t)(warp * 32) << 16) + kb * 32, hi);
      if (g.x3) {
#pragma unroll
        for (int k = 0; k < 32; ++k) w[k] = hi_of(w[k] - hi[k]);
        tm_st32(tm + ((uint32_t)(warp * 32) << 16) + kColLo + kb * 32, w);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tcf_before();
    asm volatile("bar.sync 1, 160;" ::: "memory");
    // ---- then: drain accumulators.  Thread = output column n; for each row m the warp writes 32 consecutive floats.
    const bool col_ok = n < g.Nb;
    const float bias = (g.bias && col_ok) ? __ldg(g.bias + n) : 0.f;
    uint32_t tile = 0;
    for (int64_t tb = blockIdx.x; tb < ntiles; tb += gridDim.x, ++tile) {
      const int acc = tile & 1;
      mb_wait(&tfull[acc], (tile >> 1) & 1);
      tcf_after();
      const int64_t m0 = tb * kTileRows;
#pragma unroll 1
      for (int half = 0; half < ((g.debug & 8) ? 0 : 2); ++half) {
        float v[32];
        const uint32_t ta = tm + ((uint32_t)(warp * 32) << 16) + kColAcc + acc * 128 + half * 32;
        {
          uint32_t r0[32], r1[32];
          if (g.debug & 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r0[j] = r1[j] = 0u;
          } else {
            tm_ld32_async(ta, r0);
            if (g.x3) tm_ld32_async(ta + 64, r1);
            tm_wait_ld();
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + (g.x3 ? __uint_as_float(r1[j]) : 0.f);
        }
        if (col_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int64_t m = m0 + half * 32 + j;
            if (m < g.M) {
              float o = v[j] + bias;
              if (g.accumulate) o += g.Y[m * g.ldy + n];
              if (g.pre) g.pre[m * g.ldp + n] = o;
              if (g.act == LCAO_ACT_SILU) o = o * sig_fast(o);
              if (g.G) {
                const float h = __ldg(g.G + m * g.ldg + n), sg = sig_fast(h);
                o *= sg * fmaf(h, 1.0f - sg, 1.0f);
              }
              if (!(g.debug & 1)) g.Y[m * g.ldy + n] = o;
            }
          }
        }
      }
      tcf_before();
      mb_arrive(&tempty[acc]);
    }
  }
  tcf_before();
  __syncthreads();
  if (warp == 8) {
    tcf_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

// true when LCAO_TC_TMEMW selects this kernel (default: on once validated)
bool lcao_tc_rows_t_enabled() {
  static const int on = getenv("LCAO_TC_TMEMW") ? atoi(getenv("LCAO_TC_TMEMW")) : 0;
  return on != 0;
}

int lcao_tc_rows_t(const float* A, int64_t lda, const float* W, int64_t ldw, int b_trans, const float* bias, const float* G,
                   int64_t ldg, float* Y, int64_t ldy, float* pre, int64_t ldp, int64_t M, int Kc, int Nb, int act,
                   int accumulate, int x3, cudaStream_t st) {
  RowsTArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.G = G; g.ldg = ldg; g.Y = Y; g.ldy = ldy;
  g.pre = pre; g.ldp = ldp; g.M = M; g.Kc = Kc; g.Nb = Nb; g.b_trans = b_trans; g.act = act; g.accumulate = accumulate;
  g.x3 = x3;
  static const int dbg = getenv("LCAO_TC_DEBUG") ? atoi(getenv("LCAO_TC_DEBUG")) : 0;
  g.debug = dbg;
  g.stages = x3 ? 12 : 20;
  g.lo_stages = x3 ? 12 : 1;
  const size_t smem = (size_t)(g.stages + (x3 ? g.lo_stages : 0)) * kStage + 1024 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    LCAO_CUDA(cudaFuncSetAttribute(k_tc_rows_t, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int64_t ntiles = (M + kTileRows - 1) / kTileRows;
  const unsigned grid = (unsigned)(ntiles < sm_count() ? ntiles : sm_count());
  k_tc_rows_t<<<grid, kThreadsT, smem, st>>>(g);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}
