// Three-body message passing (lcaonet.py:173-189 + shbf.py:75-87 + lcaonet.py:431-435) with the per-node dense
// products on the warp-level tensor-core path (mma.sync m16n8k8 TF32, FP32 accumulate), forward and backward.
//
// Per centre node s with in-edges i (k->s) and out-edges j (s->t), and per pair coefficients
//   a[j,i,l] = Y_l(cos(j,i)) / max(|v_ji|, eps),  |v_ji|^2 = Y^T G_i Y  (Gram form, see threebody.cu / DESIGN.md R3)
// the whole layer is three small dense products per node (GB = gate[k_i] * B[i]):
//   forward   tbw[j,:]      = sum_{i,l} a[j,i,l] GB[i,l,:]                   (dO x dI*NL) . (dI*NL x C)
//   backward  D[(i,l),j]    = sum_c GB[i,l,c] Gt[j,c]                        (dI*NL x C) . (C x dO)      Gt = d tbw
//             dGB[(i,l),:]  = sum_j a[j,i,l] Gt[j,:]                         (dI*NL x dO) . (dO x C)
// plus O(pairs) scalar work (the coefficient chain, the norm path H_i = -sum_j dot_ij a a^T, cos gradients).
// On the FP32 pipe these products were 35 of the 91 warp instructions per triplet and the kernels were issue/latency
// bound at a third of the HBM roofline (profiles/r01_notes.md).  Here they are m16n8k8 TF32 MMAs with the 3xTF32
// operand split (x = hi + lo, hi = RN to TF32; lo.hi + hi.lo + hi.hi, FP32 accumulate: FP32-equivalent, as in the dense
// layers): ~3 MMAs per pair, and the coefficient of a pair is computed by exactly the thread whose A-fragment
// needs it, so there is no shared-memory broadcast and no block barrier on the energy path.
//
// Fragment <-> data mapping (g = lane >> 2, t = lane & 3; PTX ISA m16n8k8 .tf32 layouts):
//   forward   M = 16 out-edges, K-step = (l, 8 in-edges), N-tile = 8 channels.  Thread (g,t) owns the pairs
//             (j in {g, g+8}) x (i in {t, t+4}) of a block and, for every l, builds A = a[.,.,l] from them.
//             Channel of column n in tile nt:  (nt>>2)*32 + n*4 + (nt&3)  -> every B-fragment load and every
//             output store is a float4, and a warp-wide load instruction covers 4 rows x 128 contiguous bytes.
//   backward  M = 16 rows (l, i) of a tile of IPT in-edges (IPT = 4 for NL in {3,4}, 8 for NL in {1,2}; rows g and
//             g + 8 belong to the SAME in-edge), phase 1: K = channels, N = 8 out-edges -> D; thread (g,t) then holds
//             D for the pairs (i = g % IPT) x (j in {2t, 2t+1}) whose coefficients it computes, which is exactly
//             the A-fragment of phase 2 (K = out-edges permuted k=t <-> j=2t, k=t+4 <-> j=2t+1, N = channels).
// Sums over triplets stay in accumulator registers and are written once: no atomics (forces: see kJS), deterministic.
//
// C <= 128 (one 128-channel accumulator block per warp); wider layers take the FP32-pipe kernels in threebody.cu.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tb_common.cuh"

namespace {

// x = hi + lo for the 3xTF32 product.  The tensor core reads the upper 19 bits of an operand register (sign, exponent,
// 10 mantissa bits), so hi = x truncated costs nothing — the FP32 bits are handed over as they are — and
// lo = x - trunc(x) is exact in FP32 (one LOP3 + one FADD; `cvt.rna.tf32.f32` is a 6-instruction sequence on sm_100a
// and was a third of the first version's instruction count).  lo is truncated by the tensor core in turn: the dropped
// part is below 2^-10 |lo| < 2^-20 |x|.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// d += A B in FP32-equivalent arithmetic: small terms first
__device__ __forceinline__ void mma_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                       uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d, al, bh0, bh1);
  mma_tf32(d, ah, bl0, bl1);
  mma_tf32(d, ah, bh0, bh1);
}
__device__ __forceinline__ float f4c(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// ---------------------------------------------------------------------------------------------
// per-warp shared-memory staging by bulk asynchronous copies (TMA unit, cp.async.bulk) completed on an mbarrier.
// The FP32-pipe kernels and the first tensor-core version loaded their rows with per-thread LDGs and were bound by the
// latency of each warp's dependent chain at ~12 resident warps (long-scoreboard stalls, 25 % of the DRAM bandwidth):
// a warp here puts ALL rows of its node (tens of KB) in flight with two copy instructions per in-edge and then computes
// out of shared memory; the resident warps of an SM are out of phase, so their loads overlap the others' MMAs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // (bounded: a lost copy must not hang the GPU)
  }
}
__device__ __forceinline__ float4 lds4f(const float* p) { return *reinterpret_cast<const float4*>(p); }

constexpr int kIC = 16;  // in-edges staged per chunk (forward)

// shared-memory plan of one forward warp (floats unless noted); rows are padded by 8 floats so that the four in-edges a
// quarter-warp reads in one LDS.128 fall into different bank groups
struct FwdPlan {
  int pB, pG;           // row pitches: B (NL*C + 8), gate (C + 8)
  int oB, oG, oU, oGm;  // offsets in floats
  int total;            // floats
};
__host__ __device__ inline FwdPlan fwd_plan(int C, int NL) {
  FwdPlan p;
  p.pB = NL * C + 8;
  p.pG = C + 8;
  p.oB = 4;  // [0, 16 B): the mbarrier
  p.oG = p.oB + kIC * p.pB;
  p.oU = p.oG + kIC * p.pG;
  p.oGm = p.oU + kIC * 4;                          // (8-byte aligned: every term above is a multiple of 4 floats)
  p.total = p.oGm + kIC * NL * (NL + 1);           // NP doubles per in-edge
  return p;
}

// ---------------------------------------------------------------------------------------------
// forward: CTA = one warp = one node at a time; up to 32 out-edges (two M-tiles) share every staged B fragment
// ---------------------------------------------------------------------------------------------
template <int NL, int NQ>
__global__ void __launch_bounds__(32) k_tb_fwd_mma(
    const float* __restrict__ B, int NG, const double* __restrict__ gram, const float* __restrict__ unit,
    const float* __restrict__ gate, int64_t ldg, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_edge,
    const int32_t* __restrict__ in_src, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_edge, int N,
    int C, float* __restrict__ tbw) {
  constexpr int NP = NL * (NL + 1) / 2;
  extern __shared__ __align__(16) float smem[];
  const FwdPlan pl = fwd_plan(C, NL);
  const uint32_t bar = smem_u32(smem);
  float* sB = smem + pl.oB;
  float* sG = smem + pl.oG;
  float* sU = smem + pl.oU;
  double* sGm = reinterpret_cast<double*>(smem + pl.oGm);
  const int lane = threadIdx.x, g = lane >> 2, tig = lane & 3;
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  uint32_t phase = 0;

  for (int s = blockIdx.x; s < N; s += gridDim.x) {
    const int ib = in_ptr[s], dI = in_ptr[s + 1] - ib, ob = out_ptr[s], dO = out_ptr[s + 1] - ob;
    if (dO == 0) continue;
    for (int jg = 0; jg < dO; jg += 32) {  // groups of two M-tiles (rows jg + g, + 8 | jg + 16 + g, + 24 + g)
      const int nr = min(32, dO - jg);
      const bool two = nr > 16;  // warp-uniform
      int ej[4];
      float ux[4], uy[4], uz[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        ej[r] = out_edge[ob + jg + min(g + 8 * r, nr - 1)];
        ux[r] = unit[3 * (int64_t)ej[r]]; uy[r] = unit[3 * (int64_t)ej[r] + 1]; uz[r] = unit[3 * (int64_t)ej[r] + 2];
      }
      float acc[2][NQ * 4][4];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int n = 0; n < NQ * 4; ++n)
#pragma unroll
          for (int x = 0; x < 4; ++x) acc[t][n][x] = 0.f;

      for (int ic = 0; ic < dI; ic += kIC) {
        const int n = min(kIC, dI - ic);
        // ---- stage the chunk: B rows + gate row by bulk copies (one lane per in-edge), unit / Gram through registers
        __syncwarp();  // every lane is done reading the previous chunk
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)n * (uint32_t)(NL * C + C) * 4u);
        __syncwarp();
        if (lane < n) {
          const int ep = in_edge[ib + ic + lane], k = in_src[ib + ic + lane];
          bulk_g2s(smem_u32(sB + lane * pl.pB), B + (int64_t)ep * NG * C, (uint32_t)(NL * C) * 4u, bar);
          bulk_g2s(smem_u32(sG + lane * pl.pG), gate + (int64_t)k * ldg, (uint32_t)C * 4u, bar);
          sU[lane * 4] = unit[3 * (int64_t)ep]; sU[lane * 4 + 1] = unit[3 * (int64_t)ep + 1]; sU[lane * 4 + 2] = unit[3 * (int64_t)ep + 2];
          sU[lane * 4 + 3] = __int_as_float(ep);
#pragma unroll
          for (int x = 0; x < NP; ++x) sGm[lane * NP + x] = gram[(int64_t)ep * NP + x];
        }
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1u;

        for (int i0 = 0; i0 < n; i0 += 8) {  // K-steps (l, 8 in-edges): thread (g,t) owns in-edges i0 + t, i0 + t + 4
          const int il0 = min(i0 + tig, n - 1), il1 = min(i0 + tig + 4, n - 1);
          const bool vi0 = i0 + tig < n, vi1 = i0 + tig + 4 < n;
          // ---- coefficients of this thread's pairs -> A fragments; slot (ii, jj): in-edge ii, out-edge row g + 8 jj
          uint32_t ah[2][NL][4], al[2][NL][4];
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            const int il = ii ? il1 : il0;
            const bool vi = ii ? vi1 : vi0;
            const float4 uv = lds4f(sU + il * 4);
            const int ep = __float_as_int(uv.w);
            double gm[NP];
#pragma unroll
            for (int x = 0; x < NP; ++x) gm[x] = sGm[il * NP + x];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if (r < 2 || two) {
                const float c = fmaf(ux[r], uv.x, fmaf(uy[r], uv.y, uz[r] * uv.z));
                float Y[4];
                sph_harm<NL>(c, Y);
                float w = 1.0f / fmaxf(sqrtf(fmaxf((float)quad_form<NL>(gm, Y), 0.f)), kEps);
                if (g + 8 * r >= nr || !vi || ep == ej[r]) w = 0.f;
#pragma unroll
                for (int l = 0; l < NL; ++l) split_tf32(w * Y[l], ah[r >> 1][l][ii * 2 + (r & 1)], al[r >> 1][l][ii * 2 + (r & 1)]);
              }
            }
          }
          // ---- B fragments out of shared memory: float4 at channel q*32 + g*4 (component r <-> N-tile 4q + r)
          const float* b0p = sB + il0 * pl.pB + g * 4;
          const float* b1p = sB + il1 * pl.pB + g * 4;
          const float* g0p = sG + il0 * pl.pG + g * 4;
          const float* g1p = sG + il1 * pl.pG + g * 4;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const bool ok = q * 32 + g * 4 < C;
            const float4 gt0 = ok ? lds4f(g0p + q * 32) : zero4(), gt1 = ok ? lds4f(g1p + q * 32) : zero4();
#pragma unroll
            for (int l = 0; l < NL; ++l) {
              const float4 x0 = ok ? mul4(gt0, lds4f(b0p + l * C + q * 32)) : zero4();
              const float4 x1 = ok ? mul4(gt1, lds4f(b1p + l * C + q * 32)) : zero4();
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(f4c(x0, r), bh0, bl0);
                split_tf32(f4c(x1, r), bh1, bl1);
                mma_3x(acc[0][q * 4 + r], ah[0][l], al[0][l], bh0, bh1, bl0, bl1);
                if (two) mma_3x(acc[1][q * 4 + r], ah[1][l], al[1][l], bh0, bh1, bl0, bl1);
              }
            }
          }
        }
      }
      // ---- accumulator (row g / g+8, columns 2t, 2t+1 of tile 4q+r) -> channels q*32 + 8t + r and + 4 + r
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t == 0 || two) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int ch = q * 32 + 8 * tig + 4 * hh;
              if (ch < C) {
                if (g + 16 * t < nr)
                  st4(tbw + (int64_t)ej[2 * t] * C + ch,
                      make_float4(acc[t][q * 4][hh], acc[t][q * 4 + 1][hh], acc[t][q * 4 + 2][hh], acc[t][q * 4 + 3][hh]));
                if (g + 8 + 16 * t < nr)
                  st4(tbw + (int64_t)ej[2 * t + 1] * C + ch,
                      make_float4(acc[t][q * 4][2 + hh], acc[t][q * 4 + 1][2 + hh], acc[t][q * 4 + 2][2 + hh], acc[t][q * 4 + 3][2 + hh]));
              }
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: CTA = node, warp = tile of IPT in-edges against all out-edges of the node.  Formulas as in threebody.cu:
//   dot_ij = sum_l a_l D_l ;  H_i -= [|v| > eps] dot_ij a a^T ;  dB[i,l] = gate dGB[i,l] + sum_l' H_i[l,l'] B[i,l'] + dP[i]
//   q[i] = gate (1 - gate) sum_l B[i,l] dGB[i,l]      (d xk[k] = sum over the out-edges of k of q)
//   FORCES: dcos_ij = sum_l w Y'_l D_l - [|v| > eps] w^2 dot_ij Y'^T G Y ;  d unit[e_j] += dcos unit[e_i] and vice versa
// ---------------------------------------------------------------------------------------------
constexpr int kJB = 8;  // out-edge blocks of 8 per chunk: the phase-2 A-fragments of a chunk wait in shared memory

template <int NL, int NQ, bool FORCES>
__global__ void __launch_bounds__(128, FORCES ? 2 : 3) k_tb_bwd_mma(
    const float* __restrict__ B, int NG, const double* __restrict__ gram, const float* __restrict__ unit,
    const float* __restrict__ gate, int64_t ldg, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_edge,
    const int32_t* __restrict__ in_src, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_edge, int N,
    int C, const float* __restrict__ d_tbw, const float* __restrict__ dP, float* __restrict__ dB, float* __restrict__ q,
    float* __restrict__ du_ks, float* __restrict__ du_st) {
  constexpr int NP = NL * (NL + 1) / 2;
  constexpr int NLp = NL <= 2 ? 2 : 4, IPT = 16 / NLp, HALF = NLp / 2;
  constexpr int NM = NQ * 2;  // 16-channel groups of the phase-1 contraction
  __shared__ __align__(16) float s_af[4][kJB * 32 * 4];
  __shared__ float s_st[FORCES ? 4 : 1][FORCES ? kJS * 3 : 1];  // per-warp partials of d unit[e_j] (s->t role)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int il = g & (IPT - 1), lsel = g / IPT;  // in-edge of the tile; rows g / g+8 carry l = lsel / lsel + HALF
  const int l_lo = lsel, l_hi = lsel + HALF;
  const bool vl_lo = l_lo < NL, vl_hi = l_hi < NL;
  const int NGP = NG - NL + 1;  // rows of the compact two-body gradient dP: [all l < NL | valence slot]
  float* saf = s_af[warp];

  for (int s = blockIdx.x; s < N; s += gridDim.x) {
    const int ib = in_ptr[s], dI = in_ptr[s + 1] - ib, ob = out_ptr[s], dO = out_ptr[s + 1] - ob;
    if constexpr (FORCES) {
      __syncthreads();  // the previous node's d_unit partials have been consumed
      if (dI == 0) {    // no in-edges: the s->t role gradients of the out-edges are zero
        for (int t = threadIdx.x; t < dO * 3; t += 128) du_st[3 * (int64_t)out_edge[ob + t / 3] + t % 3] = 0.f;
        continue;
      }
      for (int t = threadIdx.x; t < 4 * kJS * 3; t += 128) (&s_st[0][0])[t] = 0.f;
      for (int t = threadIdx.x; t < dO * 3; t += 128)
        if (t >= kJS * 3) du_st[3 * (int64_t)out_edge[ob + t / 3] + t % 3] = 0.f;  // rare overflow rows: global atomics below
      __syncthreads();
    }
    const int ntile = (dI + IPT - 1) / IPT;
    for (int it = warp; it < ntile; it += 4) {
      const int i = it * IPT + il;
      const bool vi = i < dI;
      const int p = ib + min(i, dI - 1);
      const int ep = in_edge[p], k = in_src[p];
      const float vx = unit[3 * (int64_t)ep], vy = unit[3 * (int64_t)ep + 1], vz = unit[3 * (int64_t)ep + 2];
      double gm[NP];
#pragma unroll
      for (int x = 0; x < NP; ++x) gm[x] = gram[(int64_t)ep * NP + x];
      const float* gate_k = gate + (int64_t)k * ldg;
      const float* B_lo = B + ((int64_t)ep * NG + (vl_lo ? l_lo : 0)) * C;
      const float* B_hi = B + ((int64_t)ep * NG + (vl_hi ? l_hi : 0)) * C;
      // phase-1 A operand: GB rows (l_lo, i) and (l_hi, i), float4 at channel m*16 + t*4 (K-steps 2m, 2m+1)
      float4 xlo[NM], xhi[NM];
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const int ch = m * 16 + tig * 4;
        const bool ok = ch < C;
        const float4 gt = ok ? ldg4(gate_k + ch) : zero4();
        xlo[m] = (ok && vl_lo) ? mul4(gt, ldg4(B_lo + ch)) : zero4();
        xhi[m] = (ok && vl_hi) ? mul4(gt, ldg4(B_hi + ch)) : zero4();
      }
      float h[NP];
#pragma unroll
      for (int x = 0; x < NP; ++x) h[x] = 0.f;
      float ks_x = 0.f, ks_y = 0.f, ks_z = 0.f;  // FORCES: d unit[e_i] partial (k->s role)

      // ---- phase 1 over one chunk of <= 64 out-edges: D = GB . Gt^T, pair scalars, phase-2 A-fragments -> saf
      auto phase1 = [&](auto from_reg, int jc) {
        constexpr bool FROMREG = decltype(from_reg)::value;
        const int njb = min(kJB, (dO - jc + 7) >> 3);
        for (int jb = 0; jb < njb; ++jb) {
          const int j0 = jc + jb * 8;
          const float* grow = d_tbw + (int64_t)out_edge[ob + min(j0 + g, dO - 1)] * C;  // B operand: out-edge j0 + g
          float D0[4] = {0.f, 0.f, 0.f, 0.f}, D1[4] = {0.f, 0.f, 0.f, 0.f}, D2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            const int ch = m * 16 + tig * 4;
            const bool ok = ch < C;
            const float4 gr = ok ? ldg4(grow + ch) : zero4();
            float4 a_lo, a_hi;
            if constexpr (FROMREG) {
              a_lo = xlo[m];
              a_hi = xhi[m];
            } else {  // later chunks of a hub node: the rows come back from L1/L2 instead of living in registers
              const float4 gt = ok ? ldg4(gate_k + ch) : zero4();
              a_lo = (ok && vl_lo) ? mul4(gt, ldg4(B_lo + ch)) : zero4();
              a_hi = (ok && vl_hi) ? mul4(gt, ldg4(B_hi + ch)) : zero4();
            }
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
              split_tf32(f4c(a_lo, 2 * h2), ah[0], al[0]);
              split_tf32(f4c(a_hi, 2 * h2), ah[1], al[1]);
              split_tf32(f4c(a_lo, 2 * h2 + 1), ah[2], al[2]);
              split_tf32(f4c(a_hi, 2 * h2 + 1), ah[3], al[3]);
              split_tf32(f4c(gr, 2 * h2), bh0, bl0);
              split_tf32(f4c(gr, 2 * h2 + 1), bh1, bl1);
              mma_tf32(D0, al, bh0, bh1);  // three independent accumulation chains (the MMA latency is ~22 cycles)
              mma_tf32(D1, ah, bl0, bl1);
              mma_tf32(D2, ah, bh0, bh1);
            }
          }
          // D[pp] = row (l_lo, i) x out-edge j0 + 2t + pp ; D[2 + pp] = row (l_hi, i) x the same out-edge
          float af[4];
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const float d_lo = (D0[pp] + D1[pp]) + D2[pp], d_hi = (D0[2 + pp] + D1[2 + pp]) + D2[2 + pp];
            const int j = j0 + 2 * tig + pp;
            const int ej = out_edge[ob + min(j, dO - 1)];
            const float ox = unit[3 * (int64_t)ej], oy = unit[3 * (int64_t)ej + 1], oz = unit[3 * (int64_t)ej + 2];
            const float cc = fmaf(ox, vx, fmaf(oy, vy, oz * vz));
            float Y[4];
            sph_harm<NL>(cc, Y);
            const float nrm = sqrtf(fmaxf((float)quad_form<NL>(gm, Y), 0.f));
            const bool live = vi && j < dO && ej != ep;
            const float ww = live ? 1.0f / fmaxf(nrm, kEps) : 0.f;
            const float fl = (live && nrm > kEps) ? 1.f : 0.f;
            float a[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) a[l] = ww * Y[l];
            const float a_lo = lsel ? a[1] : a[0], a_hi = lsel ? a[(HALF + 1) & 3] : a[HALF];
            float pd = fmaf(a_lo, d_lo, a_hi * d_hi);
            if (IPT == 4) pd += __shfl_xor_sync(0xffffffffu, pd, 16);  // the other two l of this in-edge
            const float sc = -fl * pd;
            {
              int x = 0;
#pragma unroll
              for (int la = 0; la < NL; ++la)
#pragma unroll
                for (int lb = la; lb < NL; ++lb) h[x++] += sc * a[la] * a[lb];
            }
            af[pp * 2] = a_lo;
            af[pp * 2 + 1] = a_hi;
            if constexpr (FORCES) {
              float dY[4];
              sph_harm_grad<NL>(cc, dY);
              const float y_lo = lsel ? dY[1] : dY[0], y_hi = lsel ? dY[(HALF + 1) & 3] : dY[HALF];
              float pd2 = ww * fmaf(y_lo, d_lo, y_hi * d_hi);
              if (IPT == 4) pd2 += __shfl_xor_sync(0xffffffffu, pd2, 16);
              const float corr = (float)bilin_form<NL>(gm, dY, Y);
              const float dc = live ? pd2 - fl * ww * ww * pd * corr : 0.f;
              ks_x = fmaf(dc, ox, ks_x); ks_y = fmaf(dc, oy, ks_y); ks_z = fmaf(dc, oz, ks_z);
              // d unit[e_j] += dcos unit[e_i], summed over the in-edges of the tile (lanes differing in `il`)
              float sx = dc * vx, sy = dc * vy, sz = dc * vz;
#pragma unroll
              for (int o = 4; o < 4 * IPT; o <<= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
              }
              if (g == 0 && j < dO) {
                if (j < kJS) {  // slot j of this warp's partials is only ever touched by this lane
                  s_st[warp][3 * j] += sx; s_st[warp][3 * j + 1] += sy; s_st[warp][3 * j + 2] += sz;
                } else {
                  atomicAdd(du_st + 3 * (int64_t)ej, sx);
                  atomicAdd(du_st + 3 * (int64_t)ej + 1, sy);
                  atomicAdd(du_st + 3 * (int64_t)ej + 2, sz);
                }
              }
            }
          }
          // phase-2 A fragment of this block: (row g, k=t <-> j=2t), (row g+8, k=t), (row g, k=t+4 <-> j=2t+1), (row g+8, k=t+4)
          st4(saf + (jb * 32 + lane) * 4, make_float4(af[0], af[1], af[2], af[3]));
        }
      };

      // ---- phase 2 over the same chunk: dGB += a^T . Gt
      float acc[NQ * 4][4];
      auto phase2 = [&](int jc) {
        const int njb = min(kJB, (dO - jc + 7) >> 3);
        for (int jb = 0; jb < njb; ++jb) {
          const int j0 = jc + jb * 8;
          const float4 af4 = lds4(saf + (jb * 32 + lane) * 4);  // (this lane's own slot: no cross-lane hazard)
          uint32_t ah[4], al[4];
          split_tf32(af4.x, ah[0], al[0]);
          split_tf32(af4.y, ah[1], al[1]);
          split_tf32(af4.z, ah[2], al[2]);
          split_tf32(af4.w, ah[3], al[3]);
          const float* r0 = d_tbw + (int64_t)out_edge[ob + min(j0 + 2 * tig, dO - 1)] * C;
          const float* r1 = d_tbw + (int64_t)out_edge[ob + min(j0 + 2 * tig + 1, dO - 1)] * C;
#pragma unroll
          for (int qq = 0; qq < NQ; ++qq) {
            const int ch = qq * 32 + g * 4;
            const bool ok = ch < C;
            const float4 v0 = ok ? ldg4(r0 + ch) : zero4(), v1 = ok ? ldg4(r1 + ch) : zero4();
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              uint32_t bh0, bl0, bh1, bl1;
              split_tf32(f4c(v0, r), bh0, bl0);
              split_tf32(f4c(v1, r), bh1, bl1);
              mma_3x(acc[qq * 4 + r], ah, al, bh0, bh1, bl0, bl1);
            }
          }
        }
      };

      phase1(std::true_type{}, 0);
#pragma unroll
      for (int n = 0; n < NQ * 4; ++n)
#pragma unroll
        for (int x = 0; x < 4; ++x) acc[n][x] = 0.f;
      phase2(0);
      for (int jc = kJB * 8; jc < dO; jc += kJB * 8) {  // hub nodes (> 64 out-edges)
        phase1(std::false_type{}, jc);
        phase2(jc);
      }

      // ---- finish this tile
#pragma unroll
      for (int x = 0; x < NP; ++x) {  // h was accumulated per (t, pp): add the four lanes of the in-edge's row group
        h[x] += __shfl_xor_sync(0xffffffffu, h[x], 1);
        h[x] += __shfl_xor_sync(0xffffffffu, h[x], 2);
      }
      float Hlo[NL], Hhi[NL];
      {
        float H[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) H[x][y] = 0.f;
        int x = 0;
#pragma unroll
        for (int la = 0; la < NL; ++la)
#pragma unroll
          for (int lb = la; lb < NL; ++lb) { H[la][lb] = h[x]; H[lb][la] = h[x]; ++x; }
#pragma unroll
        for (int l2 = 0; l2 < NL; ++l2) {
          Hlo[l2] = lsel ? H[1][l2] : H[0][l2];
          Hhi[l2] = lsel ? H[(HALF + 1) & 3][l2] : H[HALF][l2];
        }
      }
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int ch = qq * 32 + 8 * tig + 4 * hh;
          const bool ok = ch < C;
          const float4 gt = ok ? ldg4(gate_k + ch) : zero4();
          float4 b[NL];
#pragma unroll
          for (int l = 0; l < NL; ++l) b[l] = ok ? ldg4(B + ((int64_t)ep * NG + l) * C + ch) : zero4();
          const float4 two = (ok && dP) ? ldg4(dP + (int64_t)ep * NGP * C + ch) : zero4();  // same for every l < NL
          const float4 dlo = make_float4(acc[qq * 4][hh], acc[qq * 4 + 1][hh], acc[qq * 4 + 2][hh], acc[qq * 4 + 3][hh]);
          const float4 dhi = make_float4(acc[qq * 4][2 + hh], acc[qq * 4 + 1][2 + hh], acc[qq * 4 + 2][2 + hh], acc[qq * 4 + 3][2 + hh]);
          float4 olo = fma4(1.0f, mul4(gt, dlo), two), ohi = fma4(1.0f, mul4(gt, dhi), two);
#pragma unroll
          for (int l2 = 0; l2 < NL; ++l2) {
            olo = fma4(Hlo[l2], b[l2], olo);
            ohi = fma4(Hhi[l2], b[l2], ohi);
          }
          const float4 b_lo = lsel ? b[NL > 1 ? 1 : 0] : b[0];
          const float4 b_hi = lsel ? b[NL > HALF + 1 ? HALF + 1 : 0] : b[NL > HALF ? HALF : 0];
          float4 pq = vl_lo ? mul4(b_lo, dlo) : zero4();
          if (vl_hi) pq = add4(pq, mul4(b_hi, dhi));
          if (IPT == 4) {
            pq.x += __shfl_xor_sync(0xffffffffu, pq.x, 16);
            pq.y += __shfl_xor_sync(0xffffffffu, pq.y, 16);
            pq.z += __shfl_xor_sync(0xffffffffu, pq.z, 16);
            pq.w += __shfl_xor_sync(0xffffffffu, pq.w, 16);
          }
          if (ok && vi) {
            if (vl_lo) st4(dB + ((int64_t)ep * NG + l_lo) * C + ch, olo);
            if (vl_hi) st4(dB + ((int64_t)ep * NG + l_hi) * C + ch, ohi);
            if (lsel == 0) {
              const float4 sg = make_float4(gt.x * (1.f - gt.x), gt.y * (1.f - gt.y), gt.z * (1.f - gt.z), gt.w * (1.f - gt.w));
              st4(q + (int64_t)ep * C + ch, mul4(pq, sg));
              for (int l = NL; l < NG; ++l)  // valence slot: only the two-body path reaches it
                st4(dB + ((int64_t)ep * NG + l) * C + ch, dP ? ldg4(dP + ((int64_t)ep * NGP + 1) * C + ch) : zero4());
            }
          }
        }
      }
      if constexpr (FORCES) {
        ks_x += __shfl_xor_sync(0xffffffffu, ks_x, 1); ks_x += __shfl_xor_sync(0xffffffffu, ks_x, 2);
        ks_y += __shfl_xor_sync(0xffffffffu, ks_y, 1); ks_y += __shfl_xor_sync(0xffffffffu, ks_y, 2);
        ks_z += __shfl_xor_sync(0xffffffffu, ks_z, 1); ks_z += __shfl_xor_sync(0xffffffffu, ks_z, 2);
        if (vi && lsel == 0 && tig == 0) {
          du_ks[3 * (int64_t)ep] = ks_x; du_ks[3 * (int64_t)ep + 1] = ks_y; du_ks[3 * (int64_t)ep + 2] = ks_z;
        }
      }
    }
    if constexpr (FORCES) {
      __syncthreads();
      for (int t = threadIdx.x; t < min(dO, kJS) * 3; t += 128)
        du_st[3 * (int64_t)out_edge[ob + t / 3] + t % 3] = ((s_st[0][t] + s_st[1][t]) + s_st[2][t]) + s_st[3][t];
    }
  }
}

int grid_knob(const char* env, int dflt, int lo, int hi) {
  const char* s = getenv(env);
  const int v = s ? atoi(s) : dflt;
  return (v < lo || v > hi) ? dflt : v;
}

}  // namespace

#define TBM_DISPATCH(NL, NQ, CALL)  \
  switch ((NL) * 10 + (NQ)) {       \
    case 11: { CALL(1, 1); } break; \
    case 14: { CALL(1, 4); } break; \
    case 21: { CALL(2, 1); } break; \
    case 24: { CALL(2, 4); } break; \
    case 31: { CALL(3, 1); } break; \
    case 34: { CALL(3, 4); } break; \
    case 41: { CALL(4, 1); } break; \
    default: { CALL(4, 4); } break; \
  }

// (argument checks are done by the C entry points in threebody.cu, which dispatch here for C <= 128)
int lcao_tb_mma_fwd(const float* B, int32_t NG, const double* gram, const float* unit, const float* gate, int64_t ldg,
                    const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src, const int32_t* out_ptr,
                    const int32_t* out_edge, int64_t N, int32_t C, int32_t NL, float* tbw, cudaStream_t st) {
  static const int per_sm = grid_knob("LCAO_TBM_GRID_FWD", 24, 1, 64);
  const unsigned grid = (unsigned)(N < 148ll * per_sm ? N : 148ll * per_sm);
  const int NQ = C <= 32 ? 1 : 4;
  const size_t smem = sizeof(float) * (size_t)fwd_plan(C, NL).total;
#define CALL(nl, nq)                                                                                                   \
  {                                                                                                                    \
    static bool attr_done = false;                                                                                     \
    if (!attr_done) {                                                                                                  \
      LCAO_CUDA(cudaFuncSetAttribute(k_tb_fwd_mma<nl, nq>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));    \
      attr_done = true;                                                                                                \
    }                                                                                                                  \
    k_tb_fwd_mma<nl, nq><<<grid, 32, smem, st>>>(B, NG, gram, unit, gate, ldg, in_ptr, in_edge, in_src, out_ptr, out_edge, \
                                                 (int)N, C, tbw);                                                      \
  }
  TBM_DISPATCH(NL, NQ, CALL)
#undef CALL
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

int lcao_tb_mma_bwd(const float* B, int32_t NG, const double* gram, const float* unit, const float* gate, int64_t ldg,
                    const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src, const int32_t* out_ptr,
                    const int32_t* out_edge, int64_t N, int32_t C, int32_t NL, const float* d_tbw, const float* dP, float* dB,
                    float* q, float* du_ks, float* du_st, cudaStream_t st) {
  static const int per_sm = grid_knob("LCAO_TBM_GRID_BWD", 12, 1, 64);
  const unsigned grid = (unsigned)(N < 148ll * per_sm ? N : 148ll * per_sm);
  const int NQ = C <= 32 ? 1 : 4;
  const bool forces = du_ks != nullptr;
#define ARGS B, NG, gram, unit, gate, ldg, in_ptr, in_edge, in_src, out_ptr, out_edge, (int)N, C, d_tbw, dP, dB, q, du_ks, du_st
#define CALL(nl, nq)                                                      \
  if (forces) k_tb_bwd_mma<nl, nq, true><<<grid, 128, 0, st>>>(ARGS);     \
  else k_tb_bwd_mma<nl, nq, false><<<grid, 128, 0, st>>>(ARGS)
  TBM_DISPATCH(NL, NQ, CALL)
#undef CALL
#undef ARGS
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}
