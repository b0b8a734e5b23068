// Index construction for the LCAO hot path: by-target / by-source CSR, triplet offsets and the
// reference's triplet lists, all deterministic and bit-exact (integer work, HBM/latency bound).
// Replaces torch_sparse.SparseTensor at lcaonet.py:462-477 (argsort + CSR select + boolean mask).
#include "common.cuh"

namespace {

// warp-aggregated: lanes holding the same key elect a leader that issues one atomic for the group
// (few hot keys — species / pair tables — would otherwise serialise on a handful of addresses)
__device__ __forceinline__ int32_t grouped_atomic_add(int32_t* cnt, int64_t key, bool valid) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? key : (int64_t)-1 - lane);
  const int leader = __ffs(peers) - 1;
  int32_t base = 0;
  if (valid && (int)lane == leader) base = atomicAdd(&cnt[key], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void k_hist64(const int64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ cnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  grouped_atomic_add(cnt, i < n ? keys[i] : 0, i < n);
}

// few hot keys (species / pair tables): lanes holding the same key elect a leader that issues ONE atomic for the group
__global__ void k_hist_f(const int64_t* __restrict__ keys, int64_t n, int64_t nb, float* __restrict__ cnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const int64_t k = i < n ? keys[i] : -1;
  const bool valid = k >= 0 && k < nb;
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? k : (int64_t)-1 - lane);
  if (valid && (int)lane == __ffs(peers) - 1) atomicAdd(&cnt[k], (float)__popc(peers));  // integer-valued: exact below 2^24
}

// exclusive scan; out has n+1 entries (out[n] = total).  in may alias out.
// One CTA scans 4096 items per pass; a single CTA walking an edge-sized array pays three block barriers plus a global
// round trip per pass (124 us at E = 253 k), so arrays beyond one pass use the two-launch form below.
__global__ void __launch_bounds__(1024) k_exscan(const int32_t* in, int64_t n, int32_t* out) {
  __shared__ int32_t wsum[32];
  __shared__ int32_t total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t carry = 0;
  for (int64_t base = 0; base < n; base += 4096) {
    int64_t idx = base + (int64_t)threadIdx.x * 4;
    int32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (idx + j < n) ? in[idx + j] : 0;
    int32_t tsum = v[0] + v[1] + v[2] + v[3];
    int32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int32_t s = wsum[lane], inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      wsum[lane] = inc - s;
      if (lane == 31) total = inc;
    }
    __syncthreads();
    int32_t run = carry + wsum[w] + incl - tsum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (idx + j < n) out[idx + j] = run;
      run += v[j];
    }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

// Two-launch scan for long arrays: (1) every CTA sums its 4096-item block into bsum[b]; (2) every CTA adds up the
// sums of the blocks before it (a few hundred at most: one strided pass + a block reduction), then scans its own
// block from that carry.  bsum (ceil(n / 4096) entries) is caller-owned scratch.
constexpr int kScanBlock = 4096;

__device__ __forceinline__ int32_t block_sum_1024(int32_t v, int32_t* wsum) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) wsum[w] = v;
  __syncthreads();
  int32_t t = wsum[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;  // every thread holds the block total
}

__global__ void __launch_bounds__(1024) k_scan_block_sums(const int32_t* in, int64_t n, int32_t* bsum) {
  __shared__ int32_t wsum[32];
  const int64_t idx = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * 4;
  int32_t s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += (idx + j < n) ? in[idx + j] : 0;
  s = block_sum_1024(s, wsum);
  if (threadIdx.x == 0) bsum[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_scan_blocks(const int32_t* in, int64_t n, const int32_t* bsum, int32_t* out) {
  __shared__ int32_t wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t c = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += 1024) c += bsum[b];
  const int32_t carry = block_sum_1024(c, wsum);
  const int64_t idx = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * 4;
  int32_t v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = (idx + j < n) ? in[idx + j] : 0;
  const int32_t tsum = v[0] + v[1] + v[2] + v[3];
  int32_t incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int32_t s = wsum[lane], inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    wsum[lane] = inc - s;
  }
  __syncthreads();
  int32_t run = carry + wsum[w] + incl - tsum;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (idx + j < n) out[idx + j] = run;
    run += v[j];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) out[n] = run;  // the last thread's running sum is the total
}

// bsum: scratch of ceil(n / 4096) int32 (distinct from in / out), or nullptr for the single-CTA form
inline void launch_exscan(const int32_t* in, int64_t n, int32_t* out, int32_t* bsum, cudaStream_t st) {
  const int64_t nblk = (n + kScanBlock - 1) / kScanBlock;
  if (nblk <= 2 || !bsum) {
    k_exscan<<<1, 1024, 0, st>>>(in, n, out);
  } else {
    k_scan_block_sums<<<(unsigned)nblk, 1024, 0, st>>>(in, n, bsum);
    k_scan_blocks<<<(unsigned)nblk, 1024, 0, st>>>(in, n, bsum, out);
  }
}

__global__ void k_fill64(const int64_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ ptr,
                         int32_t* __restrict__ cursor, int32_t* __restrict__ tmp) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool valid = i < n;
  const int64_t k = valid ? keys[i] : 0;
  const int32_t off = grouped_atomic_add(cursor, k, valid);
  if (valid) tmp[ptr[k] + off] = (int32_t)i;
}

// rank-sort inside each bucket by (sec, id): deterministic whatever order the atomics produced.
// aux (nullable) receives sec[perm[j]] as int32.
__global__ void k_rank64(const int64_t* __restrict__ keys, const int64_t* __restrict__ sec, int64_t n,
                         const int32_t* __restrict__ ptr, const int32_t* __restrict__ tmp,
                         int32_t* __restrict__ perm, int32_t* __restrict__ aux) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int32_t i = tmp[p];
  const int64_t b = keys[i];
  const int64_t si = sec ? sec[i] : 0;
  const int32_t lo = ptr[b], hi = ptr[b + 1];
  int32_t rank = 0;
  for (int32_t q = lo; q < hi; ++q) {
    const int32_t j = tmp[q];
    const int64_t sj = sec ? sec[j] : 0;
    rank += (sj < si) || (sj == si && j < i);
  }
  perm[lo + rank] = i;
  if (aux) aux[lo + rank] = (int32_t)si;
}

__global__ void k_edge_prep(const int64_t* __restrict__ ei, int64_t E, int32_t* __restrict__ src32,
                            int32_t* __restrict__ dst32) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  src32[e] = (int32_t)ei[e];
  dst32[e] = (int32_t)ei[E + e];
}

__global__ void k_tri_count(const int32_t* __restrict__ src32, const int32_t* __restrict__ dst32,
                            const int32_t* __restrict__ in_ptr, int64_t E, int32_t* __restrict__ cnt) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int32_t s = src32[e];
  cnt[e] = in_ptr[s + 1] - in_ptr[s] - (dst32[e] == s ? 1 : 0);
}

// one warp per edge e: expand its triplets (e' in in(s_e), e' != e) in in-CSR order
__global__ void k_triplets_fill(const int32_t* __restrict__ src32, const int32_t* __restrict__ in_ptr,
                                const int32_t* __restrict__ in_edge, const int32_t* __restrict__ tri_ptr, int64_t E,
                                int64_t* __restrict__ tri_k, int64_t* __restrict__ e_ks, int64_t* __restrict__ e_st,
                                const float* __restrict__ unit, float* __restrict__ cos_out) {
  const int64_t e = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  const int32_t s = src32[e];
  const int32_t lo = in_ptr[s], hi = in_ptr[s + 1];
  const int64_t base = tri_ptr[e];
  const bool has_self = (tri_ptr[e + 1] - tri_ptr[e]) != (hi - lo);
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (unit) { ux = unit[3 * e]; uy = unit[3 * e + 1]; uz = unit[3 * e + 2]; }
  int32_t skipped = 0;  // number of dropped entries before the current 32-chunk (0 or 1)
  for (int32_t j0 = lo; j0 < hi; j0 += 32) {
    const int32_t j = j0 + lane;
    int32_t ep = (j < hi) ? in_edge[j] : -1;
    const bool drop = has_self && (ep == (int32_t)e);
    const unsigned dm = __ballot_sync(0xffffffffu, drop);
    const int32_t before = skipped + __popc(dm & ((1u << lane) - 1u));
    if (j < hi && !drop) {
      const int64_t o = base + (j - lo) - before;
      tri_k[o] = src32[ep];
      e_ks[o] = ep;
      e_st[o] = e;
      if (cos_out) cos_out[o] = ux * unit[3 * ep] + uy * unit[3 * ep + 1] + uz * unit[3 * ep + 2];
    }
    skipped += __popc(dm);
  }
}

// Range check of the caller's integer inputs before any of them indexes device memory (the reference raises
// IndexError from nn.Embedding / index_select for the same inputs: embed.py:41,91, base.py:38, lcaonet.py:462).
// status bits: 1 = z outside [1, max_z], 2 = batch index outside [0, n_graph), 4 = edge_index outside [0, N).
__global__ void k_validate(const int64_t* __restrict__ z, int64_t N, int64_t max_z, const int64_t* __restrict__ batch,
                           int64_t n_graph, const int64_t* __restrict__ ei, int64_t E, int32_t* __restrict__ status) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int bad = 0;
  if (i < N) {
    if (z && (z[i] < 1 || z[i] > max_z)) bad |= 1;
    if (batch && (batch[i] < 0 || batch[i] >= n_graph)) bad |= 2;
  }
  if (i < 2 * E && (ei[i] < 0 || ei[i] >= N)) bad |= 4;
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (bad && (threadIdx.x & 31) == 0) atomicOr(status, bad);
}

// ---- ORDERED grouping (stable = 2): few huge buckets (species / species-pair keys), items of a bucket in ascending id.
// Three passes over blocks of kOrdBlock consecutive items: per-block key histograms (shared-memory integer atomics:
// order independent), an exclusive scan of every key's column over the blocks (starting at the bucket's offset), and a
// fill in which one warp per block takes its 32-item groups in order and ranks equal keys inside a group with
// match_any — no position depends on the order in which atomics land, so the permutation (and every keyed reduction
// that walks it) is reproducible from run to run.
constexpr int kOrdBlock = 1024;

__global__ void __launch_bounds__(256) k_ord_hist(const int64_t* __restrict__ keys, int64_t n, int nb, int32_t* __restrict__ hist) {
  extern __shared__ int32_t s_cnt[];
  for (int k = threadIdx.x; k < nb; k += 256) s_cnt[k] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kOrdBlock;
  for (int j = threadIdx.x; j < kOrdBlock; j += 256) {
    const int64_t i = base + j;
    if (i < n) atomicAdd(&s_cnt[keys[i]], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nb; k += 256) hist[(int64_t)blockIdx.x * nb + k] = s_cnt[k];
}

__global__ void k_ord_scan(int32_t* __restrict__ hist, int nblk, int nb, const int32_t* __restrict__ ptr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nb) return;
  int32_t run = ptr[k];
  for (int b0 = 0; b0 < nblk; b0 += 8) {  // eight independent loads in flight, then the running sums in block order
    int32_t t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = (b0 + u < nblk) ? hist[(int64_t)(b0 + u) * nb + k] : 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (b0 + u < nblk) hist[(int64_t)(b0 + u) * nb + k] = run;
      run += t[u];
    }
  }
}

// one warp per block of kOrdBlock items: its 32-item groups in ascending order, keys preloaded
__global__ void __launch_bounds__(32) k_ord_fill(const int64_t* __restrict__ keys, int64_t n, int nb, const int32_t* __restrict__ hist,
                                                 int32_t* __restrict__ perm) {
  extern __shared__ int32_t s_off[];
  const int lane = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * kOrdBlock;
  int32_t key[kOrdBlock / 32];
#pragma unroll
  for (int g = 0; g < kOrdBlock / 32; ++g) {
    const int64_t i = base + g * 32 + lane;
    key[g] = i < n ? (int32_t)keys[i] : -1 - lane;  // (distinct negative keys: no peers)
  }
  for (int k = lane; k < nb; k += 32) s_off[k] = hist[(int64_t)blockIdx.x * nb + k];
  __syncwarp();
#pragma unroll
  for (int g = 0; g < kOrdBlock / 32; ++g) {
    const int32_t k = key[g];
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    if (k >= 0) perm[s_off[k] + __popc(peers & ((1u << lane) - 1u))] = (int32_t)(base + g * 32 + lane);
    __syncwarp();
    if (k >= 0 && lane == __ffs(peers) - 1) s_off[k] += __popc(peers);
    __syncwarp();
  }
}

int bucket_sort_impl(const int64_t* keys, const int64_t* sec, int64_t n, int64_t nb, int32_t* ptr, int32_t* perm,
                     int32_t* aux, int32_t* scratch, int stable, cudaStream_t st) {
  int32_t* cursor = scratch;     // nb
  int32_t* tmp = scratch + nb;   // n  (stable = 2: per-block histograms, nb * ceil(n / kOrdBlock), follow)
  LCAO_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * nb, st));
  if (n > 0) {
    k_hist64<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(keys, n, cursor);
    LCAO_LAUNCH_CHECK();
  }
  launch_exscan(cursor, nb, ptr, n >= (nb + kScanBlock - 1) / kScanBlock ? tmp : nullptr, st);  // tmp is free until the fill
  LCAO_LAUNCH_CHECK();
  if (n > 0) {
    LCAO_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * nb, st));
    if (stable == 2) {
      const int nblk = (int)ceil_div64(n, kOrdBlock);
      int32_t* hist = scratch + nb + n;
      const size_t smem = sizeof(int32_t) * (size_t)nb;
      if (smem > 48 * 1024) {
        LCAO_CUDA(cudaFuncSetAttribute(k_ord_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LCAO_CUDA(cudaFuncSetAttribute(k_ord_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      }
      k_ord_hist<<<nblk, 256, smem, st>>>(keys, n, (int)nb, hist);
      LCAO_LAUNCH_CHECK();
      k_ord_scan<<<(unsigned)ceil_div64(nb, 256), 256, 0, st>>>(hist, nblk, (int)nb, ptr);
      LCAO_LAUNCH_CHECK();
      k_ord_fill<<<nblk, 32, smem, st>>>(keys, n, (int)nb, hist, perm);
      LCAO_LAUNCH_CHECK();
      return LCAO_OK;
    }
    k_fill64<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(keys, n, ptr, cursor, stable ? tmp : perm);
    LCAO_LAUNCH_CHECK();
    if (stable) {
      k_rank64<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(keys, sec, n, ptr, tmp, perm, aux);
      LCAO_LAUNCH_CHECK();
    }
  }
  return LCAO_OK;
}

}  // namespace

extern "C" int lcao_bucket_sort(const int64_t* keys, const int64_t* sec, int64_t n, int64_t nb, int32_t* ptr,
                                int32_t* perm, int32_t* scratch, int32_t stable, void* stream) {
  LCAO_REQUIRE(n >= 0 && nb >= 0 && ptr && scratch && (n == 0 || (keys && perm)), "lcao_bucket_sort: bad arguments");
  LCAO_REQUIRE(n < (1ll << 31) && nb < (1ll << 31), "lcao_bucket_sort: sizes must fit int32");
  LCAO_REQUIRE(stable >= 0 && stable <= 2 && (stable != 2 || nb <= 50000), "lcao_bucket_sort: stable must be 0, 1 or 2 (2: at most 50000 buckets)");
  return bucket_sort_impl(keys, sec, n, nb, ptr, perm, nullptr, scratch, stable, (cudaStream_t)stream);
}

extern "C" int lcao_validate_graph(const int64_t* z, int64_t N, int64_t max_z, const int64_t* batch, int64_t n_graph,
                                   const int64_t* edge_index, int64_t E, int32_t* status, void* stream) {
  LCAO_REQUIRE(status && N >= 0 && E >= 0 && (E == 0 || edge_index), "lcao_validate_graph: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  LCAO_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  const int64_t n = N > 2 * E ? N : 2 * E;
  if (n > 0) {
    k_validate<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(z, N, max_z, batch, n_graph, edge_index, E, status);
    LCAO_LAUNCH_CHECK();
  }
  return LCAO_OK;
}

extern "C" int lcao_graph_index_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* src32, int32_t* dst32,
                                      int32_t* in_ptr, int32_t* in_edge, int32_t* in_src, int32_t* out_ptr,
                                      int32_t* out_edge, int32_t* tri_ptr, int32_t* scratch, void* stream) {
  LCAO_REQUIRE(E >= 0 && N >= 0 && in_ptr && out_ptr && scratch, "lcao_graph_index_build: null buffer");
  LCAO_REQUIRE(E == 0 || (edge_index && src32 && dst32 && in_edge && in_src && out_edge),
               "lcao_graph_index_build: null edge buffer");
  LCAO_REQUIRE(E < (1ll << 31) && N < (1ll << 31), "lcao_graph_index_build: sizes must fit int32");
  cudaStream_t st = (cudaStream_t)stream;
  // in-CSR: key = target, secondary = source ; out-CSR: key = source, secondary none
  int32_t* scr_in = scratch;               // N + E
  int32_t* scr_out = scratch + N + E;      // N + E
  int rc = bucket_sort_impl(edge_index + E, edge_index, E, N, in_ptr, in_edge, in_src, scr_in, true, st);
  if (rc) return rc;
  rc = bucket_sort_impl(edge_index, nullptr, E, N, out_ptr, out_edge, nullptr, scr_out, true, st);
  if (rc) return rc;
  if (E > 0) {
    k_edge_prep<<<(unsigned)ceil_div64(E, 256), 256, 0, st>>>(edge_index, E, src32, dst32);
    LCAO_LAUNCH_CHECK();
    if (tri_ptr) {  // only the reference's triplet LISTS need these offsets; the fused kernels work from the CSRs
      int32_t* cnt = scr_out;  // E entries, free after the out sort
      k_tri_count<<<(unsigned)ceil_div64(E, 256), 256, 0, st>>>(src32, dst32, in_ptr, E, cnt);
      LCAO_LAUNCH_CHECK();
      launch_exscan(cnt, E, tri_ptr, scr_in, st);  // scr_in (N + E entries) is free after the sorts
      LCAO_LAUNCH_CHECK();
    }
  } else if (tri_ptr) {
    LCAO_CUDA(cudaMemsetAsync(tri_ptr, 0, sizeof(int32_t), st));
  }
  return LCAO_OK;
}

// triplet offsets alone, from an existing index (for callers that asked lcao_graph_index_build to skip them)
extern "C" int lcao_triplet_offsets(const int32_t* src32, const int32_t* dst32, const int32_t* in_ptr, int64_t E,
                                    int32_t* tri_ptr, int32_t* scratch, void* stream) {
  LCAO_REQUIRE(tri_ptr && (E == 0 || (src32 && dst32 && in_ptr && scratch)), "lcao_triplet_offsets: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (E == 0) {
    LCAO_CUDA(cudaMemsetAsync(tri_ptr, 0, sizeof(int32_t), st));
    return LCAO_OK;
  }
  k_tri_count<<<(unsigned)ceil_div64(E, 256), 256, 0, st>>>(src32, dst32, in_ptr, E, scratch);
  LCAO_LAUNCH_CHECK();
  launch_exscan(scratch, E, tri_ptr, scratch + E, st);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_triplets_fill(const int32_t* src32, const int32_t* in_ptr, const int32_t* in_edge,
                                  const int32_t* tri_ptr, int64_t E, int64_t* tri_k, int64_t* e_ks, int64_t* e_st,
                                  const float* unit, float* cos_out, void* stream) {
  if (E == 0) return LCAO_OK;
  LCAO_REQUIRE(src32 && in_ptr && in_edge && tri_ptr && tri_k && e_ks && e_st, "lcao_triplets_fill: null buffer");
  LCAO_REQUIRE(!cos_out || unit, "lcao_triplets_fill: cos_out needs unit");
  k_triplets_fill<<<(unsigned)ceil_div64(E * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      src32, in_ptr, in_edge, tri_ptr, E, tri_k, e_ks, e_st, unit, cos_out);
  LCAO_LAUNCH_CHECK();
  return LCAO_OK;
}

extern "C" int lcao_histogram(const int64_t* keys, int64_t n, int64_t nb, float* counts, void* stream) {
  LCAO_REQUIRE(counts && nb > 0 && (n == 0 || keys), "lcao_histogram: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  LCAO_CUDA(cudaMemsetAsync(counts, 0, sizeof(float) * nb, st));
  if (n > 0) {
    k_hist_f<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(keys, n, nb, counts);
    LCAO_LAUNCH_CHECK();
  }
  return LCAO_OK;
}
