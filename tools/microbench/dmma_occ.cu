// DMMA m8n8k4 throughput against the number of resident warps per SM and independent accumulator chains per warp:
// what the tensor-core K2 variant (ccrs_linmma.cu: 16 warps per SM, 6 chains per warp) can expect from the FP64 pipe.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a dmma_occ.cu -o dmma_occ
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k(double* out, int iters) {
  double c[CH][2];
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run(double* out, int sms, int warps_per_sm) {
  const int iters = 1 << 12;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
  const int ctas_per_sm = warps_per_sm * 32 / threads;
  k<CH><<<sms * ctas_per_sm, threads>>>(out, iters); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<CH><<<sms * ctas_per_sm, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  const double fl = (double)sms * warps_per_sm * iters * CH * 512.0;
  printf("warps/SM %2d  chains %d : %7.2f TFLOP/s  (%.1f cycles per DMMA per sub-partition at 1.965 GHz)\n", warps_per_sm, CH,
         fl / (best * 1e-3) / 1e12, (best * 1e-3 * 1.965e9) / ((double)iters * CH * warps_per_sm / 4.0));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 2048 * 8);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  for (int w : {4, 8, 16, 32}) { run<1>(out, p.multiProcessorCount, w); run<2>(out, p.multiProcessorCount, w); run<6>(out, p.multiProcessorCount, w); run<8>(out, p.multiProcessorCount, w); }
  return 0;
}
