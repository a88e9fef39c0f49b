// scan_common.cuh -- helpers shared by the posting-list scan kernels.
#pragma once
#include "common.cuh"

#define SCAN_MAX_WARPS 32
enum { SCAN_PQ_FAST = 0, SCAN_PQ_GENERIC = 1, SCAN_FLAT_L2 = 2, SCAN_FLAT_DOT = 3 };

__device__ __forceinline__ uint4 ldg_stream16(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_stream16f(const void *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// prmt.b32 with the full selector semantics (bit 3 of a selector nibble replicates the sign bit of the chosen byte);
// the __byte_perm intrinsic only honours 3 bits per nibble.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / named barrier wrappers (warp-specialised kernels) ----------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// one lane polls (with back-off) and the warp re-converges: 32x fewer try_wait operations competing with the scan for
// the shared-memory pipe
__device__ __forceinline__ void warp_mbar_wait(uint64_t *bar, uint32_t parity, uint32_t backoff_ns = 100) {
  if ((threadIdx.x & 31) == 0) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(backoff_ns);
  }
  __syncwarp();
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
