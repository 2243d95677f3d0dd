// async_copy.cuh -- cp.async (LDGSTS) helpers shared by the MDP and POMDP
// kernels.  Copies are issued where they are written and waited for with a
// counting barrier (LDGDEPBAR / DEPBAR.LE in SASS).
#pragma once
#include <stdint.h>

namespace pp2d {

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t saddr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;"
               :: "r"(saddr), "l"(g), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
}  // namespace pp2d
