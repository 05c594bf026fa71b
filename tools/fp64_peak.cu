// Microbenchmark: sustained DFMA issue rate per SM (independent chains, full occupancy).  nvcc -arch=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> void run(int blocks_per_sm, int threads) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double *out; cudaMalloc(&out, sizeof(double) * sms * blocks_per_sm * threads);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<sms * blocks_per_sm, threads>>>(out, 100, 0.999, 0.001);
  cudaEventRecord(e0);
  k<ILP><<<sms * blocks_per_sm, threads>>>(out, iters, 0.999, 0.001);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma_total = (double)sms * blocks_per_sm * threads * iters * ILP;
  double per_sm_per_clk = fma_total / sms / (ms * 1e-3) / (clk * 1e3);
  printf("ILP=%d blocks/SM=%d threads=%d: %.2f ms, %.2f TFLOP/s FP64, %.1f DFMA lanes/clk/SM (at %d MHz nominal)\n", ILP, blocks_per_sm, threads, ms,
         2 * fma_total / (ms * 1e-3) / 1e12, per_sm_per_clk, clk / 1000);
  cudaFree(out);
}
int main() {
  run<1>(8, 256); run<2>(8, 256); run<4>(8, 256); run<8>(8, 256); run<8>(4, 256); run<4>(2, 256); run<1>(1, 128); run<1>(1, 32);
  return 0;
}
