// Warp-level tensor-core helpers of the three-body kernels: mma.sync m16n8k8 TF32 and the 3xTF32 operand split.
#pragma once
#include <stdint.h>

namespace {

// x = hi + lo for the 3xTF32 product: hi = x rounded to TF32 (10 mantissa bits) by an integer add + mask on the FP32
// bits (round half away from zero; `cvt.rna.tf32.f32` is a 6-instruction sequence on sm_100a and was a third of the first
// version's instruction count), lo = x - hi exactly.  The tensor core reads the upper 19 bits of an operand register,
// so lo is handed over as FP32 bits: the part it drops is below 2^-10 |lo| <= 2^-21 |x|.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
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

}  // namespace
