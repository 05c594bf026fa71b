#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void k_w(double *o, long n) { for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) o[i] = i; }
__global__ void k_r(const double *a, long n, double *sink) { double s = 0; for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s += a[i]; if (s == 1.2345) *sink = s; }
int main() {
  long n = 5500000;  // 44 MB
  double *hin, *hout, *din, *dout, *sink; cudaMalloc(&sink, 8);
  cudaHostAlloc(&hin, n * 8, 0); cudaHostAlloc(&hout, n * 8, 0); cudaMalloc(&din, n * 8); cudaMalloc(&dout, n * 8);
  for (long i = 0; i < n; ++i) hin[i] = i;
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); float ms;
  for (int rep = 0; rep < 2; ++rep) {
    // A: memcpy H2D + kernel writes to host
    cudaDeviceSynchronize(); cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    for (int r = 0; r < 10; ++r) { cudaMemcpyAsync(din, hin, n * 8, cudaMemcpyHostToDevice, s1); k_w<<<148, 256, 0, s2>>>(hout, n); }
    cudaEventRecord(e1, s1); cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("A memcpy H2D || kernel host-writes : %.3f ms per 44+44 MB\n", ms / 10);
    // B: memcpy D2H + kernel reads from host
    cudaDeviceSynchronize(); cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    for (int r = 0; r < 10; ++r) { cudaMemcpyAsync(hout, dout, n * 8, cudaMemcpyDeviceToHost, s1); k_r<<<148, 256, 0, s2>>>(hin, n, sink); }
    cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("B memcpy D2H || kernel host-reads  : %.3f ms\n", ms / 10);
    // C: memcpy both
    cudaDeviceSynchronize(); cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    for (int r = 0; r < 10; ++r) { cudaMemcpyAsync(din, hin, n * 8, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(hout, dout, n * 8, cudaMemcpyDeviceToHost, s2); }
    cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("C memcpy H2D || memcpy D2H         : %.3f ms\n", ms / 10);
    // D: memcpy in 1 MB pieces both directions (copy-count overhead)
    cudaDeviceSynchronize(); cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    long piece = 131072;
    for (int r = 0; r < 10; ++r) for (long o = 0; o < n; o += piece) { long m = n - o < piece ? n - o : piece; cudaMemcpyAsync(din + o, hin + o, m * 8, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(hout + o, dout + o, m * 8, cudaMemcpyDeviceToHost, s2); }
    cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("D same in 1 MB pieces (%ld copies each way): %.3f ms\n", (n + piece - 1) / piece, ms / 10);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
