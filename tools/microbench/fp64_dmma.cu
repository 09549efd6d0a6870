// Is the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) a second FP64 engine on B200, or the same one?
// Measures (a) DMMA alone, (b) DFMA alone, (c) both interleaved in the same warps, in FP64 flop/s.
// K2's Gram update could in principle be phrased as J^T J on DMMA tiles (16x16 padded from 13x13, no sparsity,
// both triangles): 3 m8n8k4 tiles x 512 flop per 2 observations' 4 rows = 3.3x the 132 DFMA it needs today, so DMMA
// would have to be >3x faster than the DFMA pipe AND run beside it to pay off.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_dmma.cu -o fp64_dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: DMMA only, 1: DFMA only, 2: both
__global__ void __launch_bounds__(256) k(double* out, int iters) {
  double c[8][2], f[8];
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; f[i] = i * 0.5; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE != 1) dmma(c[i][0], c[i][1], a, b);
      if (MODE != 0) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(double* out, int ctas, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<ctas, 256>>>(out, iters); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k<MODE><<<ctas, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  return best * 1e-3;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int ctas = p.multiProcessorCount * 8, iters = 1 << 13;
  double* out; cudaMalloc(&out, (size_t)ctas * 256 * 8);
  const double warps = (double)ctas * 8, n = (double)iters * 8;
  const double t0 = run<0>(out, ctas, iters), t1 = run<1>(out, ctas, iters), t2 = run<2>(out, ctas, iters);
  const double dmma_fl = warps * n * 512.0, dfma_fl = warps * n * 32.0 * 2.0;
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  printf("DMMA m8n8k4 alone : %.3f ms  %.2f TFLOP/s\n", t0 * 1e3, dmma_fl / t0 / 1e12);
  printf("DFMA alone        : %.3f ms  %.2f TFLOP/s\n", t1 * 1e3, dfma_fl / t1 / 1e12);
  printf("DMMA + DFMA mixed : %.3f ms  (sum of the two alone: %.3f ms, max: %.3f ms) -> %s\n", t2 * 1e3, (t0 + t1) * 1e3,
         (t0 > t1 ? t0 : t1) * 1e3, t2 > 0.85 * (t0 + t1) ? "they share one pipe" : "they overlap");
  cudaFree(out);
  return 0;
}
