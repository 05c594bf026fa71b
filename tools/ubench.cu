// tools/ubench.cu — micro-measurements that size the persistent generation kernel (profiles/r2_summary.md):
//   1. dependent DFMA latency and the issue rate with 1, 2, 4 independent chains per thread at 6 warps per scheduler
//   2. grid barrier (ticket atomic + release flag) across one CTA per resident slot (3 per SM x 256 threads)
//   3. atomicAdd on ONE address from every warp of the grid (dynamic tile counter) and int64 red.add on 20 addresses
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu ; run: tools/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int CH>
__global__ void k_dfma(double *out, int n, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-9 + c;
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_barrier(unsigned *arrive, volatile unsigned *release, int iters, long long *cycles) {
  const unsigned n = gridDim.x;
  long long t0 = clock64();
  for (int g = 0; g < iters; ++g) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned ticket = atomicAdd(arrive, 1u);
      if (ticket == (unsigned)(g + 1) * n - 1) {
        __threadfence();
        *release = (unsigned)(g + 1);
      } else {
        while (*release < (unsigned)(g + 1)) { }
      }
      __threadfence();
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

__global__ void k_atomic_one(unsigned *ctr, int per_warp, unsigned *sink) {
  unsigned acc = 0;
  for (int i = 0; i < per_warp; ++i) {
    unsigned t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(ctr, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    acc += t;
  }
  if (acc == 0xFFFFFFFFu) *sink = acc;
}

__global__ void k_red20(unsigned long long *acc20, int reps) {
  // every CTA adds 20 int64 values once per rep (the per-CTA statistics flush)
  for (int r = 0; r < reps; ++r)
    if (threadIdx.x < 20) atomicAdd(acc20 + threadIdx.x * 16, (unsigned long long)(blockIdx.x + r));
}

int main() {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
  const int sms = pr.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", pr.name, sms, pr.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 3 * 256 * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  // 1. DFMA: grid of 3 CTAs/SM x 256 threads (6 warps per scheduler), and one single warp
  const int n = 20000;
#define RUN_DFMA(CH, grid, block, tag)                                                        \
  k_dfma<CH><<<grid, block>>>(out, 100, 1.0000001, 1e-9); cudaDeviceSynchronize();            \
  cudaEventRecord(e0); k_dfma<CH><<<grid, block>>>(out, n, 1.0000001, 1e-9); cudaEventRecord(e1); \
  CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);                             \
  printf("dfma %s chains=%d: %.3f ms, %.2f cycles per dependent step, %.3f warp-DFMA/clk/SMSP\n", tag, CH, ms,      \
         ms * 1e-3 * pr.clockRate * 1e3 / n, (double)CH * n * ((grid) * (block) / 32.0) / (sms * 4.0) / (ms * 1e-3 * pr.clockRate * 1e3));
  RUN_DFMA(1, 1, 32, "1 warp");
  RUN_DFMA(2, 1, 32, "1 warp");
  RUN_DFMA(4, 1, 32, "1 warp");
  RUN_DFMA(1, sms * 3, 256, "24 warps/SM");
  RUN_DFMA(2, sms * 3, 256, "24 warps/SM");
  RUN_DFMA(4, sms * 3, 256, "24 warps/SM");
  RUN_DFMA(1, sms * 2, 256, "16 warps/SM");
  RUN_DFMA(2, sms * 2, 256, "16 warps/SM");
  RUN_DFMA(4, sms * 2, 256, "16 warps/SM");
  // 2. grid barrier
  unsigned *arrive, *release; long long *cyc;
  CK(cudaMalloc(&arrive, 256)); CK(cudaMalloc(&release, 256)); CK(cudaMalloc(&cyc, 8));
  for (int per_sm = 1; per_sm <= 3; per_sm += 2) {
    CK(cudaMemset(arrive, 0, 256)); CK(cudaMemset(release, 0, 256));
    const int iters = 2000;
    int grid = sms * per_sm;
    void *args[] = {&arrive, &release, (void *)&iters, &cyc};
    cudaEventRecord(e0);
    CK(cudaLaunchCooperativeKernel((void *)k_barrier, dim3(grid), dim3(256), args, 0, 0));
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("grid barrier, %d CTAs x 256: %.3f us per barrier\n", grid, ms * 1e3 / iters);
  }
  // 3. one-address atomics: 31250 per "sweep"
  unsigned *ctr, *sink; CK(cudaMalloc(&ctr, 256)); CK(cudaMalloc(&sink, 256));
  for (int per_warp = 9; per_warp <= 90; per_warp *= 10) {
    CK(cudaMemset(ctr, 0, 4));
    cudaEventRecord(e0); k_atomic_one<<<sms * 3, 256>>>(ctr, per_warp, sink); cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("one-address atomicAdd with return: %d atomics in %.2f us -> %.2f ns each\n", sms * 3 * 8 * per_warp, ms * 1e3, ms * 1e6 / (sms * 3 * 8 * per_warp));
  }
  unsigned long long *acc20; CK(cudaMalloc(&acc20, 20 * 16 * 8)); CK(cudaMemset(acc20, 0, 20 * 16 * 8));
  cudaEventRecord(e0); k_red20<<<sms * 3, 256>>>(acc20, 100); cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
  printf("20-address u64 atomicAdd from %d CTAs x 100 reps: %.2f us per rep\n", sms * 3, ms * 1e3 / 100);
  return 0;
}
