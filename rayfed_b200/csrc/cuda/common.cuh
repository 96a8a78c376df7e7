// Shared device helpers for the rayfed_b200 sm_100a kernels.
//
// Memory-model conventions used by every cross-GPU kernel here:
//   * payload stores to peer memory are plain (weak) vector stores;
//   * a CTA makes them visible system-wide with fence.acq_rel.sys
//     (__threadfence_system) before it signals;
//   * signals are st.release.sys / red.release.sys to a word in the consumer's
//     memory; consumers poll with ld.acquire.sys.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fedk {

constexpr int kMaxParties = 16;

struct PeerPtrs {
  void* p[kMaxParties];
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* addr, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// Streaming 16-byte accesses. Loads bypass L1 allocation (data is touched once);
// the .nc path is used only for buffers that are read-only for the kernel.
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// Coherent (non-.nc) variant for memory that a peer GPU may have just written.
__device__ __forceinline__ uint4 ld_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ uint2 ld_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_v2(void* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Nanosecond wall clock shared by every SM.  (clock64() is a *per-SM* cycle counter:
// a long-parked CTA that is context-switched and resumed elsewhere sees it jump by
// an arbitrary amount, which fired spurious timeouts on B200 — measured, see
// profiles/README.md "device-side wait".)
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spin until *flag >= target (wrap-safe signed compare). One thread per CTA
// should call this, then __syncthreads(). `timeout_ns` <= 0 waits for ever.
// Returns false on timeout.
__device__ __forceinline__ bool spin_wait_ge(const uint32_t* flag, uint32_t target, long long timeout_ns) {
  if ((int32_t)(ld_acquire_sys_u32(flag) - target) >= 0) return true;
  const unsigned long long t0 = global_ns();
  uint32_t ns = 20;
  while ((int32_t)(ld_acquire_sys_u32(flag) - target) < 0) {
    __nanosleep(ns);
    if (ns < 512) ns <<= 1;
    if (timeout_ns > 0 && (long long)(global_ns() - t0) > timeout_ns) {
      // one last look: the flag may have landed while we were asleep
      return (int32_t)(ld_acquire_sys_u32(flag) - target) >= 0;
    }
  }
  return true;
}

}  // namespace fedk
