// Asynchronous-copy plumbing shared by the staged three-body kernels: mbarriers, bulk global->shared copies (TMA unit,
// cp.async.bulk) and bounded waits.
#pragma once
#include <stdint.h>

namespace {

// per-warp shared-memory staging by bulk asynchronous copies (TMA unit, cp.async.bulk) completed on an mbarrier.
// The FP32-pipe kernels and the first tensor-core version loaded their rows with per-thread LDGs and were bound by the
// latency of each warp's dependent chain at ~12 resident warps (long-scoreboard stalls, 25 % of the DRAM bandwidth):
// a warp here puts ALL rows of its node (tens of KB) in flight with two copy instructions per in-edge and then computes
// out of shared memory; the resident warps of an SM are out of phase, so their loads overlap the others' MMAs.
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
// Bounded waits: a lost copy / protocol bug must trap instead of hanging the GPU, but a legitimate wait can be long —
// under programmatic dependent launch the producer warp sits in griddepcontrol.wait until the WHOLE previous kernel has
// finished while the consumers already poll their `full` barriers.  The bound is therefore wall-clock time (20 s).
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  unsigned long long t0 = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (++spins & 0xfffu) == 0) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 20000000000ull) __trap();
    }
  }
}
// the same with a back-off between polls: for a producer warp that is ahead of its consumers and must not take issue
// slots from them while it waits
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  unsigned long long t0 = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) {
      __nanosleep(256);
      if ((++spins & 0xffu) == 0) {
        const unsigned long long t = global_ns();
        if (t0 == 0) t0 = t;
        else if (t - t0 > 20000000000ull) __trap();
      }
    }
  }
}
__device__ __forceinline__ float4 lds4f(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace
