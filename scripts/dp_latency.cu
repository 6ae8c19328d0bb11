// Micro-benchmark (development aid): latency of dependent FP64 operations on one warp, in SM cycles.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  // dependent divisions
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) x = 1.0 / x + 0.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 1024;
  // dependent DFMA
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) x = fma(x, y, 0.25);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 1024;
  // float-seeded reciprocal + 2 Newton steps in double
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) {
    double r = (double)__frcp_rn((float)x);
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    x = r + 0.5;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 1024;
  // shared-memory dependent load chain
  __shared__ int idx[256];
  idx[threadIdx.x] = (threadIdx.x * 7 + 1) & 255;
  __syncthreads();
  int p = threadIdx.x;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) p = idx[p];
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 1024;
  // __syncthreads with 1024 threads
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 256;
  out[threadIdx.x] = x + p;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 8 * 8);
  for (int threads : {32, 1024}) {
    k<<<1, threads>>>(out, cyc, 1.3);
    cudaDeviceSynchronize();
    printf("threads %4d: div %lld cyc, dfma %lld cyc, frcp+2newton %lld cyc, lds chain %lld cyc, syncthreads %lld cyc\n", threads, cyc[0], cyc[1], cyc[2], cyc[3], cyc[4]);
  }
  return 0;
}
