// FP64 pipe microbenchmarks for B200 (sm_100a): dependent-chain latency of DFMA/DMUL/DADD, MUFU.RCP64H/RSQ64H
// sequences (1/x, rsqrt), and DFMA throughput as a function of resident warps per SM sub-partition and ILP.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_chain(double* out, int iters, double m, double c) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_div(double* out, int iters, double m) {
  double a = 1.0 + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) a = m / a;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void k_rsqrt(double* out, int iters, double m) {
  double a = 1.0 + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) a = rsqrt(a) + m;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void k_sqrt(double* out, int iters, double m) {
  double a = 1.0 + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) a = sqrt(a) + m;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

template <class F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 32 * 1024 * 8);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz * 1e-6;
  const int iters = 1 << 16;
  printf("clock %.3f GHz (attr)\n", ghz);
  // latency: 1 warp per SM, one chain
  float ms = time_it([&] { k_chain<1><<<148, 32>>>(out, iters, 1.0000001, 1e-9); });
  printf("DFMA dependent latency: %.2f cycles\n", ms * 1e-3 * ghz * 1e9 / iters);
  ms = time_it([&] { k_div<<<148, 32>>>(out, iters, 1.5); });
  printf("double division dependent latency: %.1f cycles\n", ms * 1e-3 * ghz * 1e9 / iters);
  ms = time_it([&] { k_rsqrt<<<148, 32>>>(out, iters, 0.5); });
  printf("rsqrt(+DADD) dependent latency: %.1f cycles\n", ms * 1e-3 * ghz * 1e9 / iters);
  ms = time_it([&] { k_sqrt<<<148, 32>>>(out, iters, 0.5); });
  printf("sqrt(+DADD) dependent latency: %.1f cycles\n", ms * 1e-3 * ghz * 1e9 / iters);
  // throughput: warps per SMSP x ILP  (cycles per warp-DFMA per SMSP)
  for (int wps : {1, 2, 3, 4, 8}) {
    const int threads = wps * 4 * 32;
    float t1 = time_it([&] { k_chain<1><<<148, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t2 = time_it([&] { k_chain<2><<<148, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t4 = time_it([&] { k_chain<4><<<148, threads>>>(out, iters, 1.0000001, 1e-9); });
    float t8 = time_it([&] { k_chain<8><<<148, threads>>>(out, iters, 1.0000001, 1e-9); });
    auto cyc = [&](float ms, int ilp) { return ms * 1e-3 * ghz * 1e9 / ((double)iters * ilp * wps); };
    printf("warps/SMSP=%d  cycles per warp-DFMA per SMSP: ILP1 %.2f  ILP2 %.2f  ILP4 %.2f  ILP8 %.2f\n", wps, cyc(t1, 1), cyc(t2, 2), cyc(t4, 4), cyc(t8, 8));
  }
  return 0;
}
