// ccrs_devutil.cuh — device helpers shared by ccrs_kernels.cu and ccrs_loop.cu: self-validating result slots, the
// cross-GPU exchange over peer memory, cp.async wrappers.
#pragma once
#include "ccrs_device.cuh"
#include "ccrs_kernels.cuh"

namespace ccrs {

// bit pattern that arms a result slot which validates itself (device partials, mapped host results): a NaN payload
// no computation produces (the Cholesky-failure poison is the canonical quiet NaN)
constexpr long long kArmBits = 0x7ff8dead5e471e15LL;
static __global__ void k_arm(double* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = __longlong_as_double(kArmBits);
}

// value v of this rank's partial -> sum over ranks (see PeerXchg). Called by one thread per value.
CCRS_D double peer_exchange(const PeerXchg& px, int v, double mine) {
  const size_t slot = (size_t)px.off + v;
  for (int r = 0; r < px.world; ++r)
    *reinterpret_cast<volatile double*>(px.peer[r] + slot + (size_t)px.rank * kXchgMaxVals) = mine;
  volatile double* loc = px.peer[px.rank] + slot;
  double tot = 0.0;
  const long long t0 = clock64();
  for (int r = 0; r < px.world; ++r) {
    double x = loc[(size_t)r * kXchgMaxVals];
    while (__double_as_longlong(x) == kArmBits) {
      if (clock64() - t0 > 4000000000LL) { x = nan(""); break; }   // ~2 s: a peer never arrived -> poison, not a hang
      x = loc[(size_t)r * kXchgMaxVals];
    }
    tot += x;                                                       // rank order
    loc[(size_t)r * kXchgMaxVals] = __longlong_as_double(kArmBits); // re-arm for the next use of this area
  }
  return tot;
}

// Loads for spin loops on self-validating slots: volatile (the compiler must re-issue them every pass — a plain
// __ldcg may legally be hoisted out of the loop, which turns "not there yet" into an endless spin) and strong at GPU
// scope (served by L2, the coherence point).
CCRS_D double ld_spin(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
CCRS_D double2 ld_spin2(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}

// globaltimer (ns), low 40 bits: exact in a double, wraps every 18 minutes
CCRS_D double stamp_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (double)(t & ((1ull << 40) - 1));
}

CCRS_D void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
CCRS_D void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
// no "memory" clobber on the issue side: ordinary loads may be scheduled across the prefetch (the pose prologue's loads
// then overlap it); the wait below is the barrier that orders the ring reads
CCRS_D void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
CCRS_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace ccrs
