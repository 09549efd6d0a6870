// DFMA issue rate with realistic operand patterns: acc[i][j] += a[i] * b[j] (3 distinct 64-bit register operands),
// the shape of the Gram-block accumulation in K2. Compared with the best-case fma(a, const, const) chain.
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int NJ>
__global__ void k_outer(double* out, const double* in, int iters) {
  double a[NI], b[NJ], acc[NI][NJ];
#pragma unroll
  for (int i = 0; i < NI; ++i) a[i] = in[i] + threadIdx.x;
#pragma unroll
  for (int j = 0; j < NJ; ++j) b[j] = in[NI + j] - threadIdx.x;
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    // perturb the operands so the compiler cannot hoist anything (2 cheap ops per iteration)
    a[0] += 1e-9; b[0] -= 1e-9;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// symmetric rank-1 update of a packed upper triangle: acc[i<=j] += v[i]*v[j]  (exactly K2's pattern, N=13 -> 91 FMAs)
template <int N>
__global__ void k_syr(double* out, const double* in, int iters) {
  double v[N], acc[N * (N + 1) / 2];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = in[i] + threadIdx.x;
#pragma unroll
  for (int i = 0; i < N * (N + 1) / 2; ++i) acc[i] = 0.0;
  for (int it = 0; it < iters; ++it) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) { acc[k] = fma(v[i], v[j], acc[k]); ++k; }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += 1e-9;   // N extra DADDs per iteration
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < N * (N + 1) / 2; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double *out, *in;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&in, 64 * 8);
  cudaMemset(in, 0, 64 * 8);
  const double ghz = 1.965;
  const int iters = 1 << 14;
  for (int wps : {1, 2, 4}) {
    const int threads = wps * 128;
    float t88 = time_it([&] { k_outer<8, 8><<<148, threads>>>(out, in, iters); });
    float t412 = time_it([&] { k_outer<4, 12><<<148, threads>>>(out, in, iters); });
    float t13 = time_it([&] { k_syr<13><<<148, threads>>>(out, in, iters); });
    auto cyc = [&](float ms, int n) { return ms * 1e-3 * ghz * 1e9 / ((double)iters * n * wps); };
    printf("warps/SMSP=%d cycles per warp-DFMA per SMSP: outer8x8 %.2f (66 ops)  outer4x12 %.2f (50 ops)  syr13 %.2f (104 ops)\n", wps,
           cyc(t88, 66), cyc(t412, 50), cyc(t13, 104));
  }
  return 0;
}
