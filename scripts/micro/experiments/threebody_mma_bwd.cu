// Three-body backward (autograd of lcaonet.py:173-189 + shbf.py:75-87; algebra in threebody.cu) with both per-node dense
// products on the warp-level tensor cores (mma.sync m16n8k8 TF32, 3xTF32 operand split = FP32-equivalent) and every
// operand staged in shared memory by bulk asynchronous copies (TMA unit) issued by a producer warp.
//
// Per tile of <= 16 in-edges r (k->s) of a node and chunk of <= 16 of its out-edges j (s->t), GB = gate * B, Gt = d tbw:
//   phase 1   D_l[r, j]    = sum_c GB_l[r, c] Gt[j, c]          M = 16 in-edges, N = 8 out-edges x 2, K = channels
//   scalars   a_l[r, j]    = w Y_l(cos),  dot = sum_l a_l D_l,  H_r -= flag dot a a^T          one pair per thread slot
//   phase 3   dGB_l[r, c] += sum_j a_l[r, j] Gt[j, c]           M = 16 in-edges, N = channels, K = 8 out-edges x 2
//   epilogue  dB[r, l] = gate dGB_l + two-body + sum_l' H[l, l'] B[r, l'],  q[r] = gate (1 - gate) sum_l B_l dGB_l
// The FP32-pipe formulations (threebody.cu: 91 warp instructions per triplet; threebody_staged.cu: ~60) are bound by
// instruction issue at 0.33-0.38 of the HBM roofline; here a triplet costs ~1.5 tensor instructions plus ~20 others.
//
// CTA = 4 consumer warps + 1 producer warp.  Consumer warp w owns channels [32w, 32w + 32): in phase 1 it contracts its
// 32 channels (K split; the four partial D are summed in a fixed order through shared memory), computes a quarter of the
// pair scalars, and in phase 3 / the epilogue it owns its channels of all 16 rows (accumulators in registers across the
// chunks of a tile).  Fragment <-> data maps (g = lane >> 2, t = lane & 3):
//   phase 1  k-step ks of a warp's 32 channels:  k = t <-> channel 8t + 2ks,  k = t + 4 <-> channel 8t + 2ks + 1
//            (thread t reads 8 consecutive channels of rows g, g + 8 (A) and of out-edges g, g + 8 (B) as float4s)
//   phase 3  n-tile nt, column n <-> channel 4n + nt  (B fragments and accumulator rows are float4s)
// Dead rows / out-edges of a tile carry zero coefficients; their operands are whatever finite data the buffers hold.
// C % 32 == 0, C <= 128, energy path only (no d unit): other shapes and the forces variant take threebody_staged.cu.
#include <stdlib.h>

#include "common.cuh"
#include "tb_common.cuh"
#include "tb_async.cuh"
#include "tb_mma.cuh"

namespace {

constexpr int kTI = 16;   // in-edges per tile (MMA M)
constexpr int kTJ = 16;   // out-edges per chunk
constexpr int kGtB = 2;   // resident d_tbw chunks
constexpr int kNst = 2;   // stage ring depth (one tile being multiplied, one in flight)
constexpr int kDpJ = kTI * 4 + 4;  // floats per (l, j) row of the partial-D exchange: [r][warp], padded
enum { mFirst = 1, mLast = 2, mEnd = 4 };

struct MPlan {
  int slotp, oMeta, oGram, oItem, stage, oGt, gtp, gtbuf, oDp, oAf, oHs, total;
};
__host__ __device__ inline MPlan mma_plan(int C, int NL) {
  MPlan p;
  const int NP = NL * (NL + 1) / 2;
  p.slotp = (NL + 1) * C + 4;          // B rows | gate row, padded: rows g, g + 1 fall into different banks
  p.oMeta = kTI * p.slotp;             // kTI x float4 (unit vector, edge id)
  p.oGram = p.oMeta + kTI * 4;         // kTI x NP doubles
  p.oItem = p.oGram + kTI * NP * 2;    // int4 {nI, nO, flags, gt buffer | parity << 1}, int4 {dO, 0, 0, 0}
  p.stage = p.oItem + 8;
  p.oGt = 32 + kNst * p.stage;
  p.gtp = C + 4;
  p.gtbuf = kTJ * p.gtp + kTJ * 4;     // d_tbw rows | float4 (unit vector, edge id) per out-edge
  p.oDp = p.oGt + kGtB * p.gtbuf;      // partial D: [l][j][r][warp]
  p.oAf = p.oDp + NL * kTJ * kDpJ;     // A fragments of phase 3: [l][k-step][hi | lo][lane][4]
  p.oHs = p.oAf + NL * 2 * 2 * 128;    // norm-path partials: [r][warp][12]
  p.total = p.oHs + kTI * 48;
  return p;
}

constexpr int kCW = 8;    // consumer warps
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float f4c(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

template <int NL, int CT>
__global__ void __launch_bounds__(288, 2) k_tb_bwd_mma(
    const float* __restrict__ B, int NG, const double* __restrict__ gram, const float* __restrict__ unit,
    const float* __restrict__ gate, int64_t ldg, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_edge,
    const int32_t* __restrict__ in_src, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_edge, int N,
    int C_rt, const float* __restrict__ d_tbw, const float* __restrict__ dP, float* __restrict__ dB, float* __restrict__ q) {
  constexpr int NP = NL * (NL + 1) / 2;
  const int C = CT ? CT : C_rt;
  const int NGP = NG - NL + 1;
  extern __shared__ __align__(128) float smem[];
  const MPlan pl = mma_plan(C, NL);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(smem), bar_empty = bar_full + 32, bar_gt = bar_full + 64;
  for (int i = 32 + threadIdx.x; i < pl.total; i += blockDim.x) smem[i] = 0.f;  // dead operands must be finite
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNst; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, kCW); }
    for (int i = 0; i < kGtB; ++i) mbar_init(bar_gt + 8 * i, 1);
  }
  fence_proxy_async();
  __syncthreads();
  const int per_cta = (N + (int)gridDim.x - 1) / (int)gridDim.x;
  const int s_end = min(N, ((int)blockIdx.x + 1) * per_cta);

  if (warp == kCW) {
    // ------------------------------------------------------------------ producer (see threebody_staged.cu)
    // metadata per PANEL = (node, <= 32 of its in-edges: lane <-> in-edge; its first 32 out-edges: lane <-> out-edge),
    // fetched one panel ahead; items = (tile of 16 in-edges of the panel, chunk of 16 out-edges)
    struct Node { int s, ib, dI, ob, dO; bool valid; };
    struct Ids { int ep, k, ej; };
    struct Dat { float vx, vy, vz, ux, uy, uz; double gm[NP]; };
    auto next_node = [&](Node& n) {
      n.valid = false;
      while (n.s < s_end) {
        n.ib = in_ptr[n.s]; n.dI = in_ptr[n.s + 1] - n.ib; n.ob = out_ptr[n.s]; n.dO = out_ptr[n.s + 1] - n.ob;
        if (n.dI > 0) { n.valid = true; return; }
        n.s += 1;
      }
    };
    auto next_panel = [&](Node& n, int& ic) {
      if (!n.valid) return;
      ic += 32;
      if (ic < n.dI) return;
      ic = 0;
      n.s += 1;
      next_node(n);
    };
    auto load_ids = [&](const Node& n, int ic, Ids& r) {
      r.ep = 0; r.k = 0; r.ej = 0;
      if (!n.valid) return;
      if (ic + lane < n.dI) { r.ep = in_edge[n.ib + ic + lane]; r.k = in_src[n.ib + ic + lane]; }
      if (lane < n.dO) r.ej = out_edge[n.ob + lane];
    };
    auto load_dat = [&](const Node& n, int ic, const Ids& r, Dat& d) {
      if (!n.valid) return;
      if (ic + lane < n.dI) {
        d.vx = unit[3 * (int64_t)r.ep]; d.vy = unit[3 * (int64_t)r.ep + 1]; d.vz = unit[3 * (int64_t)r.ep + 2];
#pragma unroll
        for (int x = 0; x < NP; ++x) d.gm[x] = gram[(int64_t)r.ep * NP + x];
      }
      if (lane < n.dO) { d.ux = unit[3 * (int64_t)r.ej]; d.uy = unit[3 * (int64_t)r.ej + 1]; d.uz = unit[3 * (int64_t)r.ej + 2]; }
    };
    int it = 0, st = 0, ph = 0;
    int tagS0 = -1, tagJ0 = 0, cnt0 = 0, use0 = -1, ust0 = 0, uph0 = 0, tagS1 = -1, tagJ1 = 0, cnt1 = 0, use1 = -1, ust1 = 0, uph1 = 0;
    auto stage_wait = [&]() {
      if (it >= kNst) mbar_wait_backoff(bar_empty + 8 * st, (uint32_t)(ph ^ 1));
    };
    auto stage_next = [&]() {
      it += 1;
      if (++st == kNst) { st = 0; ph ^= 1; }
    };
    auto issue = [&](const Node& n, int ic, const Ids& r, const Dat& d, int b, int jc) {
      stage_wait();
      float* sS = smem + 32 + st * pl.stage;
      const uint32_t bar = bar_full + 8 * st;
      const int nIp = min(32, n.dI - ic), cnt = min(kTI, nIp - kTI * b), nO = max(0, min(kTJ, n.dO - jc));
      int bsel = 0, par = 0;
      if (nO > 0) {
        if (tagS0 == n.s && tagJ0 == jc) bsel = 0;
        else if (tagS1 == n.s && tagJ1 == jc) bsel = 1;
        else {
          bsel = (use0 <= use1) ? 0 : 1;  // least recently used
          const int x = bsel ? use1 : use0;
          if (x >= 0 && x > it - kNst) mbar_wait_backoff(bar_empty + 8 * (bsel ? ust1 : ust0), (uint32_t)(bsel ? uph1 : uph0));
          float* gbuf = smem + pl.oGt + bsel * pl.gtbuf;
          // rows of the chunk: out-edges jc .. jc + nO; the first 32 of a node sit in the panel registers (lane <-> out-edge)
          const int row = lane - (jc < 32 ? jc : 0);
          const bool rowok = row >= 0 && row < nO;
          int ej = r.ej;
          float ux = d.ux, uy = d.uy, uz = d.uz;
          if (jc >= 32 && rowok) {  // later chunks of a wide node: fetched on demand
            ej = out_edge[n.ob + jc + row];
            ux = unit[3 * (int64_t)ej]; uy = unit[3 * (int64_t)ej + 1]; uz = unit[3 * (int64_t)ej + 2];
          }
          if (rowok) st4(gbuf + kTJ * pl.gtp + row * 4, make_float4(ux, uy, uz, __int_as_float(ej)));
          __syncwarp();
          if (lane == 0) mbar_expect_tx(bar_gt + 8 * bsel, (uint32_t)nO * (uint32_t)C * 4u);
          __syncwarp();
          if (rowok) bulk_g2s(smem_u32(gbuf + row * pl.gtp), d_tbw + (int64_t)ej * C, (uint32_t)C * 4u, bar_gt + 8 * bsel);
          if (bsel) { tagS1 = n.s; tagJ1 = jc; cnt1 += 1; } else { tagS0 = n.s; tagJ0 = jc; cnt0 += 1; }
        }
        par = ((bsel ? cnt1 : cnt0) - 1) & 1;
        if (bsel) { use1 = it; ust1 = st; uph1 = ph; } else { use0 = it; ust0 = st; uph0 = ph; }
      }
      const bool first = jc == 0, last = jc + kTJ >= n.dO;
      const bool mine = (lane >> 4) == b && lane < nIp;  // lane <-> in-edge ic + lane -> row lane & 15 of tile lane >> 4
      const int slot = lane & 15;
      if (mine) {
        st4(sS + pl.oMeta + slot * 4, make_float4(d.vx, d.vy, d.vz, __int_as_float(r.ep)));
        double* sG = reinterpret_cast<double*>(sS + pl.oGram);
#pragma unroll
        for (int x = 0; x < NP; ++x) sG[slot * NP + x] = d.gm[x];
      }
      if (lane == 0) {
        *reinterpret_cast<int4*>(sS + pl.oItem) = make_int4(cnt, nO, (first ? mFirst : 0) | (last ? mLast : 0), bsel | (par << 1));
        *reinterpret_cast<int4*>(sS + pl.oItem + 4) = make_int4(n.dO, 0, 0, 0);
      }
      __syncwarp();
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * (uint32_t)((NL + 1) * C) * 4u);
      __syncwarp();
      if (mine) {
        bulk_g2s(smem_u32(sS + slot * pl.slotp), B + (int64_t)r.ep * NG * C, (uint32_t)(NL * C) * 4u, bar);
        bulk_g2s(smem_u32(sS + slot * pl.slotp + NL * C), gate + (int64_t)r.k * ldg, (uint32_t)C * 4u, bar);
      }
      stage_next();
    };
    Node cur, nxt;
    int cic = 0, nic = 0;
    Ids ci, ni;
    Dat cd, nd;
    cur.s = (int)blockIdx.x * per_cta;
    next_node(cur);
    load_ids(cur, cic, ci);
    load_dat(cur, cic, ci, cd);
    while (cur.valid) {
      nxt = cur; nic = cic;
      next_panel(nxt, nic);
      load_ids(nxt, nic, ni);
      const int nIp = min(32, cur.dI - cic);
      bool pending = true;
      for (int b = 0; b * kTI < nIp; ++b) {
        int jc = 0;
        do {
          issue(cur, cic, ci, cd, b, jc);
          if (pending) { load_dat(nxt, nic, ni, nd); pending = false; }
          jc += kTJ;
        } while (jc < cur.dO);
      }
      cur = nxt; cic = nic; ci = ni; cd = nd;
    }
    stage_wait();
    if (lane == 0) {
      *reinterpret_cast<int4*>(smem + 32 + st * pl.stage + pl.oItem) = make_int4(0, 0, mEnd, 0);
      mbar_expect_tx(bar_full + 8 * st, 0u);
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  // warp w: phase 1 = channels [32 (w >> 1), + 32) x out-edge n-tile (w & 1); pair scalars = one pair per thread;
  // phase 3 / epilogue = channels [16 w, + 16) of all 16 rows
  const int g = lane >> 2, t = lane & 3;
  const int q1 = (warp >> 1) * 32, nt1 = warp & 1;
  const int q3 = warp * 16;
  const bool has1 = q1 < C, has3 = q3 < C;
  float* Dp = smem + pl.oDp;
  float* Af = smem + pl.oAf;
  float* Hs = smem + pl.oHs;
  float dacc[NL][2][4];
#pragma unroll
  for (int l = 0; l < NL; ++l)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int x = 0; x < 4; ++x) dacc[l][nt][x] = 0.f;

  for (int st = 0, ph = 0;; ph ^= (st + 1 == kNst), st = (st + 1 == kNst) ? 0 : st + 1) {
    const float* sS = smem + 32 + st * pl.stage;
    mbar_wait(bar_full + 8 * st, (uint32_t)ph);
    const int4 item = *reinterpret_cast<const int4*>(sS + pl.oItem);
    const int nI = item.x, nO = item.y, flags = item.z;
    if (flags & mEnd) break;
    const int dO = (*reinterpret_cast<const int4*>(sS + pl.oItem + 4)).x;
    if (flags & mFirst) {
#pragma unroll
      for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int x = 0; x < 4; ++x) dacc[l][nt][x] = 0.f;
    }
    if (nO > 0) {
      const int gi = item.w;
      mbar_wait(bar_gt + 8 * (gi & 1), (uint32_t)(gi >> 1));
      const float* gtb = smem + pl.oGt + (gi & 1) * pl.gtbuf;
      const bool two = nO > 8;  // second n-tile (phase 1) / k-step (phase 3) is live
      // ---- phase 1: partial D_l[r, 8 nt1 + n] over the channels [q1, q1 + 32)
      if (nt1 == 0 || two) {
        float accD[NL][4];
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
          for (int x = 0; x < 4; ++x) accD[l][x] = 0.f;
        if (has1) {
          const float* r0 = sS + g * pl.slotp + q1 + 8 * t;
          const float* r1 = sS + (g + 8) * pl.slotp + q1 + 8 * t;
          const float* jp = gtb + (8 * nt1 + g) * pl.gtp + q1 + 8 * t;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {  // channels 8t + 4hf .. + 3 <-> k-steps 2hf, 2hf + 1
            const float4 gt0 = lds4f(r0 + NL * C + 4 * hf), gt1 = lds4f(r1 + NL * C + 4 * hf);
            const float4 x0 = lds4f(jp + 4 * hf);
            uint32_t bh[4], bl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32(f4c(x0, e), bh[e], bl[e]);
#pragma unroll
            for (int l = 0; l < NL; ++l) {
              const float4 y0 = mul4(gt0, lds4f(r0 + l * C + 4 * hf)), y1 = mul4(gt1, lds4f(r1 + l * C + 4 * hf));
#pragma unroll
              for (int u = 0; u < 2; ++u) {  // k-step 2hf + u: k = t <-> element 2u, k = t + 4 <-> element 2u + 1
                uint32_t ah[4], al[4];
                split_tf32(f4c(y0, 2 * u), ah[0], al[0]);
                split_tf32(f4c(y1, 2 * u), ah[1], al[1]);
                split_tf32(f4c(y0, 2 * u + 1), ah[2], al[2]);
                split_tf32(f4c(y1, 2 * u + 1), ah[3], al[3]);
                mma_3x(accD[l], ah, al, bh[2 * u], bh[2 * u + 1], bl[2 * u], bl[2 * u + 1]);
              }
            }
          }
        }
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int r = g + 8 * (x >> 1), j = 8 * nt1 + 2 * t + (x & 1);
            Dp[(l * kTJ + j) * kDpJ + r * 4 + (warp >> 1)] = accD[l][x];
          }
      }
      cbar();
      // ---- pair scalars: pair (row r, out-edge jl) = element 2 half + rr of the A fragment (lane, k-step kj)
      {
        const int kj = warp >> 2, half = (warp >> 1) & 1, rr = warp & 1;
        const int jl = 8 * kj + t + 4 * half, r = g + 8 * rr;
        const float4 om = lds4f(gtb + kTJ * pl.gtp + jl * 4);
        const float4 vm = lds4f(sS + pl.oMeta + r * 4);
        const bool live = r < nI && jl < nO && __float_as_int(om.w) != __float_as_int(vm.w);
        double gm[NP];
        {
          const double* sG = reinterpret_cast<const double*>(sS + pl.oGram) + r * NP;
#pragma unroll
          for (int p = 0; p < NP; ++p) gm[p] = sG[p];
        }
        const float cc = fmaf(om.x, vm.x, fmaf(om.y, vm.y, om.z * vm.z));
        float Y[4];
        sph_harm<NL>(cc, Y);
        const float nrm = sqrtf(fmaxf((float)quad_form<NL>(gm, Y), 0.f));
        const float ww = live ? 1.0f / fmaxf(nrm, kEps) : 0.f;
        const float fl = (live && nrm > kEps) ? 1.f : 0.f;
        float a[NL], dot = 0.f;
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          a[l] = live ? ww * Y[l] : 0.f;
          const float4 p4 = lds4f(Dp + (l * kTJ + jl) * kDpJ + r * 4);
          const float D = live ? ((p4.x + p4.y) + p4.z) + p4.w : 0.f;
          dot = fmaf(a[l], D, dot);
          uint32_t hi, lo;
          split_tf32(a[l], hi, lo);
          float* af = Af + ((l * 2 + kj) * 2) * 128 + lane * 4 + 2 * half + rr;
          af[0] = __uint_as_float(hi);
          af[128] = __uint_as_float(lo);
        }
        const float sc = -fl * dot;
        float hv[NP];
        {
          int p = 0;
#pragma unroll
          for (int x = 0; x < NL; ++x)
#pragma unroll
            for (int y = x; y < NL; ++y) hv[p++] = sc * a[x] * a[y];
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          hv[p] += __shfl_xor_sync(0xffffffffu, hv[p], 1);
          hv[p] += __shfl_xor_sync(0xffffffffu, hv[p], 2);
        }
        if (t == 0) {  // (r, warp >> 1) is owned by this lane for the whole tile
          float* hs = Hs + r * 48 + (warp >> 1) * 12;
#pragma unroll
          for (int p = 0; p < NP; ++p) hs[p] = (flags & mFirst) ? hv[p] : hs[p] + hv[p];
        }
      }
      cbar();
      // ---- phase 3: dGB_l[r, c] += sum_j a_l[r, j] Gt[j, c] on the channels [q3, q3 + 16): column n of n-tile nt <-> q3 + 2n + nt
      if (has3) {
#pragma unroll
        for (int kj = 0; kj < 2; ++kj) {
          if (kj == 0 || two) {
            const float2 za = *reinterpret_cast<const float2*>(gtb + (8 * kj + t) * pl.gtp + q3 + 2 * g);
            const float2 zb = *reinterpret_cast<const float2*>(gtb + (8 * kj + t + 4) * pl.gtp + q3 + 2 * g);
            uint32_t bh0[2], bl0[2], bh1[2], bl1[2];
            split_tf32(za.x, bh0[0], bl0[0]); split_tf32(za.y, bh0[1], bl0[1]);
            split_tf32(zb.x, bh1[0], bl1[0]); split_tf32(zb.y, bh1[1], bl1[1]);
#pragma unroll
            for (int l = 0; l < NL; ++l) {
              const float4 a_hi = lds4f(Af + ((l * 2 + kj) * 2) * 128 + lane * 4);
              const float4 a_lo = lds4f(Af + ((l * 2 + kj) * 2 + 1) * 128 + lane * 4);
              const uint32_t ah[4] = {__float_as_uint(a_hi.x), __float_as_uint(a_hi.y), __float_as_uint(a_hi.z), __float_as_uint(a_hi.w)};
              const uint32_t al[4] = {__float_as_uint(a_lo.x), __float_as_uint(a_lo.y), __float_as_uint(a_lo.z), __float_as_uint(a_lo.w)};
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) mma_3x(dacc[l][nt], ah, al, bh0[nt], bh1[nt], bl0[nt], bl1[nt]);
            }
          }
        }
      }
    }
    if ((flags & mLast) && has3) {
      // ---- epilogue: rows g, g + 8; channels q3 + 4t .. + 3 = accumulator elements (2 rr, 2 rr + 1) of n-tiles (0, 1)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = g + 8 * rr;
        if (r < nI) {
          const int ep = __float_as_int(sS[pl.oMeta + r * 4 + 3]);
          float hsum[12];
#pragma unroll
          for (int p = 0; p < 12; ++p) hsum[p] = 0.f;
          if (dO > 0) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
#pragma unroll
              for (int v = 0; v < (NP + 3) / 4; ++v) {
                const float4 hq = lds4f(Hs + r * 48 + w * 12 + 4 * v);
                hsum[4 * v] += hq.x; hsum[4 * v + 1] += hq.y; hsum[4 * v + 2] += hq.z; hsum[4 * v + 3] += hq.w;
              }
            }
          }
          float H[NL][NL];
          {
            int p = 0;
#pragma unroll
            for (int x = 0; x < NL; ++x)
#pragma unroll
              for (int y = x; y < NL; ++y) { H[x][y] = hsum[p]; H[y][x] = hsum[p]; ++p; }
          }
          const float* row = sS + r * pl.slotp;
          const int c = q3 + 4 * t;
          const float4 gtv = lds4f(row + NL * C + c);
          const float4 two_body = dP ? ldg4(dP + (int64_t)ep * NGP * C + c) : zero4();
          float4 b[NL], da[NL];
#pragma unroll
          for (int l = 0; l < NL; ++l) {
            b[l] = lds4f(row + l * C + c);
            da[l] = make_float4(dacc[l][0][2 * rr], dacc[l][1][2 * rr], dacc[l][0][2 * rr + 1], dacc[l][1][2 * rr + 1]);
          }
          float4 qq = zero4();
#pragma unroll
          for (int l = 0; l < NL; ++l) {
            float4 o4 = fma4(1.0f, mul4(gtv, da[l]), two_body);
#pragma unroll
            for (int l2 = 0; l2 < NL; ++l2) o4 = fma4(H[l][l2], b[l2], o4);
            st4(dB + ((int64_t)ep * NG + l) * C + c, o4);
            qq = add4(qq, mul4(b[l], da[l]));
          }
          qq = mul4(qq, make_float4(gtv.x * (1.f - gtv.x), gtv.y * (1.f - gtv.y), gtv.z * (1.f - gtv.z), gtv.w * (1.f - gtv.w)));
          st4(q + (int64_t)ep * C + c, qq);
          for (int l = NL; l < NG; ++l)
            st4(dB + ((int64_t)ep * NG + l) * C + c, dP ? ldg4(dP + ((int64_t)ep * NGP + 1) * C + c) : zero4());
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * st);
  }
}

int mknob(const char* env, int dflt, int lo, int hi) {
  const char* s = getenv(env);
  const int v = s ? atoi(s) : dflt;
  return (v < lo || v > hi) ? dflt : v;
}

template <int NL, int CT>
int launch_mma_bwd(const float* B, int NG, const double* gram, const float* unit, const float* gate, int64_t ldg,
                   const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src, const int32_t* out_ptr,
                   const int32_t* out_edge, int64_t N, int C, const float* d_tbw, const float* dP, float* dB, float* q,
                   cudaStream_t st) {
  static const int per_sm = mknob("LCAO_TBM_GRID_BWD", 8, 1, 128);
  const size_t smem = sizeof(float) * (size_t)mma_plan(C, NL).total;
  static bool attr_done = false;
  if (!attr_done) {
    LCAO_CUDA(cudaFuncSetAttribute(k_tb_bwd_mma<NL, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_done = true;
  }
  const int64_t want = (N + 1) / 2;
  const unsigned grid = (unsigned)(want < 148ll * per_sm ? (want > 0 ? want : 1) : 148ll * per_sm);
  k_tb_bwd_mma<NL, CT><<<grid, 288, smem, st>>>(B, NG, gram, unit, gate, ldg, in_ptr, in_edge, in_src, out_ptr, out_edge,
                                               (int)N, C, d_tbw, dP, dB, q);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

}  // namespace

// energy-path backward for C % 32 == 0, C <= 128 (dispatch and argument checks: lcao_threebody_bwd in threebody.cu)
int lcao_tb_mma_bwd(const float* B, int32_t NG, const double* gram, const float* unit, const float* gate, int64_t ldg,
                    const int32_t* in_ptr, const int32_t* in_edge, const int32_t* in_src, const int32_t* out_ptr,
                    const int32_t* out_edge, int64_t N, int32_t C, int32_t NL, const float* d_tbw, const float* dP,
                    float* dB, float* q, cudaStream_t st) {
#define ARGS B, NG, gram, unit, gate, ldg, in_ptr, in_edge, in_src, out_ptr, out_edge, N, C, d_tbw, dP, dB, q, st
#define CALL(nl) return C == 128 ? launch_mma_bwd<nl, 128>(ARGS) : launch_mma_bwd<nl, 0>(ARGS);
  switch (NL) {
    case 1: CALL(1)
    case 2: CALL(2)
    case 3: CALL(3)
    default: CALL(4)
  }
#undef CALL
#undef ARGS
}
