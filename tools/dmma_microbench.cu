// Microbenchmark: latency and throughput of DMMA.8x8x4 (mma.sync m8n8k4 f64) on sm_100a as a function of the number of
// independent accumulator chains per warp and warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CHAINS>
__global__ void k(int iters, double a, double b, double* out, long long* cyc) {
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void kfma(int iters, double a, double b, double* out, long long* cyc, int chains) {
  double c0 = threadIdx.x, c1 = 1, c2 = 2, c3 = 3;
  long long t0 = clock64();
  if (chains == 1) for (int it = 0; it < iters; ++it) c0 = fma(c0, a, b);
  else for (int it = 0; it < iters; ++it) { c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b); }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS> void run(int warps, double* out, long long* cyc) {
  const int iters = 4096;
  k<CHAINS><<<148, warps * 32>>>(iters, 1e-9, 1e-9, out, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CHAINS><<<148, warps * 32>>>(iters, 1e-9, 1e-9, out, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double n = (double)iters * CHAINS;
  printf("chains %d warps/SM %2d: %7.1f cycles per DMMA step (per warp, all chains) -> %6.1f cyc/DMMA/warp, SM rate 1 DMMA per %5.2f cyc, %6.2f TFLOP/s\n",
         CHAINS, warps, (double)c / iters, (double)c / n, (double)c / (n * warps), 148.0 * warps * n * 512 / (ms * 1e-3) / 1e12);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int w : {1, 2, 4, 8, 16, 32}) { run<1>(w, out, cyc); }
  for (int w : {1, 4, 16}) { run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc); }
  for (int ch : {1, 4}) {
    kfma<<<148, 32>>>(4096, 1.0000001, 1e-9, out, cyc, ch); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA chains %d: %.1f cycles per iteration\n", ch, (double)c / 4096);
  }
  return 0;
}
