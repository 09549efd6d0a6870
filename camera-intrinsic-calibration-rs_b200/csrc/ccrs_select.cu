// ccrs_select.cu — order statistics of the per-observation reprojection errors without sorting.
//
// The reference's validation (src/util.rs:721-795) sorts every reprojection error of a camera and reports
// sorted[n/2] (median) and the mean of the smallest n*99/100 errors. On the device a full sort is unnecessary:
// both numbers follow from two order statistics and one masked sum. Errors are non-negative doubles, so their bit
// patterns order like unsigned integers: a most-significant-digit radix select (11-bit digits, six passes over the
// array, 8 B/observation per pass, HBM-bound) finds the exact key at a given rank for two ranks at once; a last pass
// sums the values below the second key in a fixed order. Integer atomics only: bitwise reproducible.
#include "ccrs_kernels.cuh"

namespace ccrs {

constexpr int kSelThreads = 256;

// histogram of the current digit over the keys that match each target's prefix so far
__global__ void __launch_bounds__(kSelThreads) k_select_hist(const double* __restrict__ v, int64_t n, int shift, int bits,
                                                             const SelectState* __restrict__ st, unsigned* __restrict__ hist) {
  __shared__ unsigned sh[2][kSelBins];
  for (int i = threadIdx.x; i < 2 * kSelBins; i += kSelThreads) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const unsigned long long pa = st->prefix[0], pb = st->prefix[1];
  const bool first = shift + bits >= 64;           // no prefix yet: every key matches
  const unsigned mask = (1u << bits) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSelThreads) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(v[i]);
    const unsigned long long hi = first ? 0ull : (key >> (shift + bits));
    const unsigned d = (unsigned)(key >> shift) & mask;
    if (first || hi == pa) atomicAdd(&sh[0][d], 1u);
    if (first || hi == pb) atomicAdd(&sh[1][d], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kSelBins; i += kSelThreads) {
    const unsigned c = (&sh[0][0])[i];
    if (c) atomicAdd(hist + i, c);
  }
}

// one CTA of two warps: warp t advances target t by one digit, then the histogram is cleared for the next pass
__global__ void __launch_bounds__(64) k_select_scan(unsigned* __restrict__ hist, int bits, SelectState* __restrict__ st) {
  const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = 1 << bits, per = nb / 32;
  const unsigned* h = hist + t * kSelBins;
  unsigned long long mine = 0;
  for (int i = 0; i < per; ++i) mine += h[lane * per + i];
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const unsigned long long excl = incl - mine;
  const unsigned long long rank = st->rank[t];
  if (rank >= excl && rank < incl) {   // exactly one lane (the rank is below the number of matching keys)
    unsigned long long below = excl;
    int b = lane * per;
    for (; b < lane * per + per; ++b) {
      const unsigned c = h[b];
      if (rank < below + c) break;
      below += c;
    }
    st->prefix[t] = (st->prefix[t] << bits) | (unsigned long long)b;
    st->rank[t] = rank - below;
    st->below[t] += below;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kSelBins; i += 64) hist[i] = 0u;
}

// partial[c] = sum of the values whose key is below target 1's key, CTA c owning a contiguous chunk (fixed order)
__global__ void __launch_bounds__(kSelThreads) k_select_sum(const double* __restrict__ v, int64_t n,
                                                            const SelectState* __restrict__ st, double* __restrict__ partial) {
  __shared__ double sh[kSelThreads];
  const unsigned long long T = st->prefix[1];
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b = (int64_t)blockIdx.x * chunk, e = b + chunk < n ? b + chunk : n;
  double s = 0.0;
  for (int64_t i = b + threadIdx.x; i < e; i += kSelThreads) {
    const double x = v[i];
    if ((unsigned long long)__double_as_longlong(x) < T) s += x;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = kSelThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

cudaError_t launch_select(const double* v, int64_t n, SelectState* st, unsigned* hist, double* partial, int n_ctas,
                          cudaStream_t s, int64_t* launches) {
  // digits from the most significant end: 11 11 11 11 11 9 bits
  static const int widths[6] = {11, 11, 11, 11, 11, 9};
  int shift = 64;
  for (int pass = 0; pass < 6; ++pass) {
    const int bits = widths[pass];
    shift -= bits;
    k_select_hist<<<n_ctas, kSelThreads, 0, s>>>(v, n, shift, bits, st, hist);
    k_select_scan<<<1, 64, 0, s>>>(hist, bits, st);
    if (launches) *launches += 2;
  }
  k_select_sum<<<n_ctas, kSelThreads, 0, s>>>(v, n, st, partial);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace ccrs
