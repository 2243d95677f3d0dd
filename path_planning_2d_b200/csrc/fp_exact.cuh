// fp_exact.cuh -- float operations with the rounding / flush-to-zero behaviour
// of the reference's device code (nvcc --use_fast_math: FFMA.FTZ, FMUL.FTZ,
// FADD.FTZ).  Host-side reference arithmetic (x86-64, no FMA, subnormals kept)
// is reproduced with the __fmul_rn / __fadd_rn / __fdiv_rn intrinsics instead.
#pragma once

namespace pp2d {

__device__ __forceinline__ float fma_ftz(float a, float b, float c) {
  float d;
  asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float mul_ftz(float a, float b) {
  float d;
  asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float add_ftz(float a, float b) {
  float d;
  asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

}  // namespace pp2d
