// Microbenchmark of the 8x8 diagonal-tile factorisation variants (one warp), cycles per tile.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsqrt_fast(double x) {
  // float seed (MUFU.RSQ) + two Newton steps in FP64; x must be a normal positive number inside the float range
  double r = (double)rsqrtf((float)x);
  double h = 0.5 * x;
  r = r * (1.5 - h * r * r);
  r = r * (1.5 - h * r * r);
  return r;
}
template <int VARIANT>
__global__ void k(int iters, const double* A, double* out, long long* cyc) {
  __shared__ double accP[64], sW[64];
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < 64; i += 32) accP[i] = A[i];
  __syncwarp();
  long long t0 = clock64();
  double sink = 0;
  for (int it = 0; it < iters; ++it) {
    const int i = lane & 7;
    double a[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = accP[i * 8 + j] + sink * 1e-30; w[j] = (j == i) ? 1.0 : 0.0; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double akk = __shfl_sync(0xffffffffu, a[k], k);
      if (!(akk > 0.0)) akk = 1.0;
      const double r = (VARIANT == 0) ? rsqrt(akk) : (VARIANT == 1 ? rsqrt_fast(akk) : 1.0 / sqrt(akk));
      const double lik = ((i == k) ? akk : a[k]) * r;
      a[k] = lik;
      double pj[8], wj[8];
#pragma unroll
      for (int j = k + 1; j < 8; ++j) pj[j] = __shfl_sync(0xffffffffu, lik, j);
#pragma unroll
      for (int j = 0; j <= k; ++j) wj[j] = __shfl_sync(0xffffffffu, w[j], k) * r;
      if (i > k) {
#pragma unroll
        for (int j = k + 1; j < 8; ++j) a[j] -= lik * pj[j];
#pragma unroll
        for (int j = 0; j <= k; ++j) w[j] -= lik * wj[j];
      } else if (i == k) {
#pragma unroll
        for (int j = 0; j <= k; ++j) w[j] = wj[j];
      }
    }
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sW[i * 8 + j] = (j <= i) ? w[j] : 0.0;
    }
    __syncwarp();
    sink += sW[lane];
  }
  long long t1 = clock64();
  out[lane] = sink;
  if (lane == 0) *cyc = t1 - t0;
}
int main() {
  double hA[64];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) hA[i * 8 + j] = (i == j) ? 10.0 + i : 0.3 / (1 + abs(i - j));
  double *A, *out; long long* cyc;
  cudaMalloc(&A, 512); cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  cudaMemcpy(A, hA, 512, cudaMemcpyHostToDevice);
  const int iters = 2000;
  long long c;
  k<0><<<1, 32>>>(iters, A, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("GJ + rsqrt()      : %.0f cycles per tile\n", (double)c / iters);
  k<1><<<1, 32>>>(iters, A, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("GJ + rsqrt_fast   : %.0f cycles per tile\n", (double)c / iters);
  k<2><<<1, 32>>>(iters, A, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("GJ + 1/sqrt       : %.0f cycles per tile\n", (double)c / iters);
  double ho[32]; cudaMemcpy(ho, out, 256, cudaMemcpyDeviceToHost); printf("check %g\n", ho[3]);
  return 0;
}
