#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
struct P { double *a[4]; int *b[3]; double *oa[4]; int *ob[3]; };
__global__ void k_rw(P p, long n, int work) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double x0 = p.a[0][i], x1 = p.a[1][i], x2 = p.a[2][i], x3 = p.a[3][i]; int y0 = p.b[0][i], y1 = p.b[1][i], y2 = p.b[2][i];
    for (int k = 0; k < work; ++k) x0 = fma(x0, 1.0000001, x1 * 1e-9);
    p.oa[0][i] = x0 + 1; p.oa[1][i] = x1; p.oa[2][i] = x2; p.oa[3][i] = x3; p.ob[0][i] = y0; p.ob[1][i] = y1; p.ob[2][i] = y2;
  }
}
__global__ void k_r(P p, long n, double *sink) {
  double s = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    s += p.a[0][i] + p.a[1][i] + p.a[2][i] + p.a[3][i] + p.b[0][i] + p.b[1][i] + p.b[2][i];
  if (s == 1.2345) *sink = s;
}
__global__ void k_w(P p, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    p.oa[0][i] = i; p.oa[1][i] = i; p.oa[2][i] = i; p.oa[3][i] = i; p.ob[0][i] = i; p.ob[1][i] = i; p.ob[2][i] = i; }
}
int main() {
  long n = 1000000; P p; double *sink; cudaMalloc(&sink, 8);
  for (int j = 0; j < 4; ++j) { cudaHostAlloc(&p.a[j], n * 8, 0); cudaHostAlloc(&p.oa[j], n * 8, 0); for (long i = 0; i < n; ++i) p.a[j][i] = i; }
  for (int j = 0; j < 3; ++j) { cudaHostAlloc(&p.b[j], n * 4, 0); cudaHostAlloc(&p.ob[j], n * 4, 0); for (long i = 0; i < n; ++i) p.b[j][i] = i; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  int grids[] = {148, 444, 1184, 3908}; int blocks[] = {256, 1024};
  for (int b : blocks) for (int g : grids) {
    k_rw<<<g, b>>>(p, n, 0); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int r = 0; r < 10; ++r) k_rw<<<g, b>>>(p, n, 0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    float rw = ms / 10;
    cudaEventRecord(e0); for (int r = 0; r < 10; ++r) k_r<<<g, b>>>(p, n, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    float rd = ms / 10;
    cudaEventRecord(e0); for (int r = 0; r < 10; ++r) k_w<<<g, b>>>(p, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("grid %4d x %4d: rw %.3f ms (%.1f GB/s sum)  read-only %.3f ms (%.1f GB/s)  write-only %.3f ms (%.1f GB/s)\n", g, b, rw, 88e-3 * n / 1e6 / rw * 1e3 / 1e3, rd, 44e-3*n/1e6 / rd, ms / 10, 44e-3*n/1e6 / (ms / 10));
  }
  k_rw<<<444, 256>>>(p, n, 20000); cudaDeviceSynchronize();
  cudaEventRecord(e0); k_rw<<<444, 256>>>(p, n, 20000); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  printf("with ~20k dependent fma per gene: %.3f ms\n", ms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
