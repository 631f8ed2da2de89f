// Fine-scale neighbourhood solves of the local corrector problems.
//
// The reference solves  (sum_q theta_q(mu) A_q^nbh) c = f^nbh  on the neighbourhood of a marked subdomain with dune-istl
// through `lhs.apply_inverse(rhs, inverse_options)` (discretize_elliptic_block_swipdg.py:227-316, called from
// reductor.py:75-78).  Here: Jacobi-preconditioned conjugate gradients on a CSR matrix resident in HBM.  All scalars
// (alpha, beta, the inner products) stay on the device -- every CTA re-derives them from per-CTA partial sums in a fixed
// order, so an iteration needs no host round trip and the result is bit-reproducible.  kPcgCheck iterations run inside ONE
// cooperative launch (`pcg_iterate_kernel`: two phases per iteration separated by grid-wide barriers, every CTA
// keeping its rows of the matrix warm in its L1); the host looks at the residual norm between launches only.  Devices or
// contexts without cooperative launch get the same arithmetic as three launches per iteration.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kPcgThreads = 256;
constexpr int kPcgCheck = 25;
constexpr int kLanesPerRow = 4;

struct PcgVecs {
  double *r, *z, *p, *Ap, *dinv, *part_a, *part_b, *sc;   // sc: [0], [1] rz (by iteration parity), [2] rr, [3] bb, [4] breakdown flag
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // fixed-order tree over the block
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = kPcgThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double out = sh[0];
  __syncthreads();
  return out;
}

// every CTA sums all per-CTA partials in the same order -> the same value everywhere
__device__ __forceinline__ double all_partials(const double* __restrict__ part, int n_part, double* sh) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n_part; i += kPcgThreads) v += __ldcg(part + i);
  return block_sum(v, sh);
}

__device__ __forceinline__ double row_dot(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                          const double* __restrict__ values, const double* __restrict__ x, int row, int sub) {
  double s = 0.0;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  for (int p = p0 + sub; p < p1; p += kLanesPerRow) s += values[p] * __ldg(x + colind[p]);
#pragma unroll
  for (int o = kLanesPerRow / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// r = b - A x, dinv = 1 / diag(A), z = dinv r, p = z; partials of r.z (part_a) and r.r (part_b); bb
__global__ void __launch_bounds__(kPcgThreads) pcg_init_kernel(int n, const int32_t* __restrict__ rowptr,
                                                               const int32_t* __restrict__ colind,
                                                               const double* __restrict__ values, const double* __restrict__ b,
                                                               const double* __restrict__ x, PcgVecs V, double* part_bb) {
  __shared__ double sh[kPcgThreads];
  const int rows_per_block = kPcgThreads / kLanesPerRow;
  const int sub = threadIdx.x % kLanesPerRow;
  double rz = 0.0, rr = 0.0, bb = 0.0;
  for (int row0 = blockIdx.x * rows_per_block; row0 < n; row0 += gridDim.x * rows_per_block) {
    const int row = row0 + threadIdx.x / kLanesPerRow;
    const bool row_ok = row < n;
    {
      const double ax = row_dot(rowptr, colind, values, x, row_ok ? row : 0, sub);   // all lanes take part in the shuffles
      if (row_ok && sub == 0) {
        double d = 1.0;
        for (int p = rowptr[row]; p < rowptr[row + 1]; ++p)
          if (colind[p] == row) d = values[p];
        const double di = (d != 0.0) ? 1.0 / d : 1.0;
        const double rv = b[row] - ax, zv = di * rv;
        V.dinv[row] = di; V.r[row] = rv; V.z[row] = zv; V.p[row] = zv;
        rz += rv * zv; rr += rv * rv; bb += b[row] * b[row];
      }
    }
  }
  const double s0 = block_sum(rz, sh), s1 = block_sum(rr, sh), s2 = block_sum(bb, sh);
  if (threadIdx.x == 0) { V.part_a[blockIdx.x] = s0; V.part_b[blockIdx.x] = s1; part_bb[blockIdx.x] = s2; }
}

__global__ void __launch_bounds__(kPcgThreads) pcg_init_scalars_kernel(int n_part, PcgVecs V, const double* part_bb) {
  __shared__ double sh[kPcgThreads];
  const double rz = all_partials(V.part_a, n_part, sh), rr = all_partials(V.part_b, n_part, sh), bb = all_partials(part_bb, n_part, sh);
  if (threadIdx.x == 0) { V.sc[0] = rz; V.sc[1] = rz; V.sc[2] = rr; V.sc[3] = bb; V.sc[4] = 0.0; }
}

// Ap = A p; partials of p.Ap (part_a)
__global__ void __launch_bounds__(kPcgThreads) pcg_spmv_kernel(int n, const int32_t* __restrict__ rowptr,
                                                               const int32_t* __restrict__ colind,
                                                               const double* __restrict__ values, PcgVecs V) {
  __shared__ double sh[kPcgThreads];
  const int rows_per_block = kPcgThreads / kLanesPerRow;
  const int sub = threadIdx.x % kLanesPerRow;
  double pap = 0.0;
  for (int row0 = blockIdx.x * rows_per_block; row0 < n; row0 += gridDim.x * rows_per_block) {
    const int row = row0 + threadIdx.x / kLanesPerRow;
    const bool row_ok = row < n;
    const double ap = row_dot(rowptr, colind, values, V.p, row_ok ? row : 0, sub);     // all lanes take part in the shuffles
    if (row_ok && sub == 0) { V.Ap[row] = ap; pap += V.p[row] * ap; }
  }
  const double s = block_sum(pap, sh);
  if (threadIdx.x == 0) V.part_a[blockIdx.x] = s;
}

// alpha = rz / p.Ap;  x += alpha p;  r -= alpha Ap;  z = dinv r;  partials of r.z (part_b) and r.r (part_c)
__global__ void __launch_bounds__(kPcgThreads) pcg_update_xr_kernel(int n, int n_part, int parity, double* __restrict__ x, PcgVecs V,
                                                                    double* __restrict__ part_c) {
  __shared__ double sh[kPcgThreads];
  const double pap = all_partials(V.part_a, n_part, sh);
  const double rz = V.sc[parity];
  const bool ok = pap > 0.0 && V.sc[4] == 0.0;
  const double alpha = ok ? rz / pap : 0.0;
  double s_rz = 0.0, s_rr = 0.0;
  for (int i = blockIdx.x * kPcgThreads + threadIdx.x; i < n; i += gridDim.x * kPcgThreads) {
    x[i] += alpha * V.p[i];
    const double rv = V.r[i] - alpha * V.Ap[i];
    const double zv = V.dinv[i] * rv;
    V.r[i] = rv; V.z[i] = zv;
    s_rz += rv * zv; s_rr += rv * rv;
  }
  const double a = block_sum(s_rz, sh), b = block_sum(s_rr, sh);
  if (threadIdx.x == 0) { V.part_b[blockIdx.x] = a; part_c[blockIdx.x] = b; }
  if (!ok && blockIdx.x == 0 && threadIdx.x == 0 && rz != 0.0) V.sc[4] = 1.0;   // p.Ap <= 0: not positive definite
}

// beta = rz_new / rz;  p = z + beta p;  CTA 0 publishes rz_new (other parity) and r.r
__global__ void __launch_bounds__(kPcgThreads) pcg_update_p_kernel(int n, int n_part, int parity, PcgVecs V,
                                                                   const double* __restrict__ part_c) {
  __shared__ double sh[kPcgThreads];
  const double rz_new = all_partials(V.part_b, n_part, sh);
  const double rr = all_partials(part_c, n_part, sh);
  const double rz = V.sc[parity];
  const double beta = (rz != 0.0) ? rz_new / rz : 0.0;
  for (int i = blockIdx.x * kPcgThreads + threadIdx.x; i < n; i += gridDim.x * kPcgThreads) V.p[i] = V.z[i] + beta * V.p[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) { V.sc[parity ^ 1] = rz_new; V.sc[2] = rr; }
}

// ---- one cooperative launch = n_it iterations: at most one CTA of 1024 threads per SM (a grid-wide barrier costs by the
// number of CTAs), the phases of an iteration separated by grid-wide barriers.  Everything another CTA wrote is read
// through L2 (__ldcg): the L1 of this SM may still hold the previous iteration's line.  The matrix itself is read-only
// (__ldg, L1-resident across iterations: a CTA always takes the same rows).  Same recurrences as the three kernels above;
// the sums are formed in a different (equally fixed) order: shuffle tree per warp, 32 warp sums, one value per CTA.
constexpr int kCoopThreads = 1024;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the CTA, the same value in every thread; sh: 32 doubles
__device__ __forceinline__ double coop_block_sum(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();                                       // sh may still be read from the previous call
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return warp_sum(sh[threadIdx.x & 31]);
}

// every CTA sums all per-CTA partials (n_part <= 1024) in the same order -> the same value everywhere
__device__ __forceinline__ double coop_all_partials(const double* part, int n_part, double* sh) {
  return coop_block_sum(((int)threadIdx.x < n_part) ? __ldcg(part + threadIdx.x) : 0.0, sh);
}

// Two grid-wide barriers per iteration: the direction update p <- z + beta p is folded into the matrix-vector product
// (every lane forms the entries of the new p it needs from z and the old p -- the same expression as the owner of the row,
// so the same bits -- and the owner stores them into the OTHER p buffer, which nobody reads in this phase).
__global__ void __launch_bounds__(kCoopThreads, 1) pcg_iterate_kernel(int n, const int32_t* __restrict__ rowptr,
                                                                      const int32_t* __restrict__ colind,
                                                                      const double* __restrict__ values, double* x, PcgVecs V,
                                                                      double* p_alt, double* part_c, int it0, int n_it) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[32];
  const int n_part = gridDim.x;
  const int rows_per_block = kCoopThreads / kLanesPerRow;
  const int sub = threadIdx.x % kLanesPerRow;
  double rz_cur = __ldcg(V.sc + (it0 & 1)), rz_prev = __ldcg(V.sc + ((it0 & 1) ^ 1));
  bool broken = __ldcg(V.sc + 4) != 0.0;
  for (int k = 0; k < n_it; ++k) {
    const int it = it0 + k;
    const double* p_old = (it & 1) ? p_alt : V.p;
    double* p_new = (it & 1) ? V.p : p_alt;
    const double beta = (it == 0 || rz_prev == 0.0) ? 0.0 : rz_cur / rz_prev;
    // ---- p_new = z + beta p_old (own rows stored);  Ap = A p_new;  partials of p_new.Ap
    {
      double pap = 0.0;
      for (int row0 = blockIdx.x * rows_per_block; row0 < n; row0 += gridDim.x * rows_per_block) {
        const int row = row0 + threadIdx.x / kLanesPerRow;
        const bool row_ok = row < n;
        const int r = row_ok ? row : 0;
        double s = 0.0;
        const int q0 = __ldg(rowptr + r), q1 = __ldg(rowptr + r + 1);
        for (int q = q0 + sub; q < q1; q += kLanesPerRow) {
          const int c = __ldg(colind + q);
          s += __ldg(values + q) * (__ldcg(V.z + c) + beta * __ldcg(p_old + c));
        }
#pragma unroll
        for (int o = kLanesPerRow / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row_ok && sub == 0) {
          const double pv = __ldcg(V.z + row) + beta * __ldcg(p_old + row);
          __stcg(p_new + row, pv);
          __stcg(V.Ap + row, s);
          pap += pv * s;
        }
      }
      const double sum = coop_block_sum(pap, sh);
      if (threadIdx.x == 0) __stcg(V.part_a + blockIdx.x, sum);
    }
    grid.sync();
    // ---- alpha = rz / p.Ap;  x += alpha p;  r -= alpha Ap;  z = dinv r;  partials of r.z and r.r
    {
      const double pap = coop_all_partials(V.part_a, n_part, sh);
      const bool ok = pap > 0.0 && !broken;
      const double alpha = ok ? rz_cur / pap : 0.0;
      if (!ok && rz_cur != 0.0) broken = true;                // p.Ap <= 0: not positive definite (the same in every CTA)
      double s_rz = 0.0, s_rr = 0.0;
      for (int i = blockIdx.x * kCoopThreads + threadIdx.x; i < n; i += gridDim.x * kCoopThreads) {
        __stcg(x + i, __ldcg(x + i) + alpha * __ldcg(p_new + i));
        const double rv = __ldcg(V.r + i) - alpha * __ldcg(V.Ap + i);
        const double zv = __ldg(V.dinv + i) * rv;
        __stcg(V.r + i, rv);
        __stcg(V.z + i, zv);
        s_rz += rv * zv; s_rr += rv * rv;
      }
      const double a = coop_block_sum(s_rz, sh), b = coop_block_sum(s_rr, sh);
      if (threadIdx.x == 0) { __stcg(V.part_b + blockIdx.x, a); __stcg(part_c + blockIdx.x, b); }
    }
    grid.sync();
    rz_prev = rz_cur;
    rz_cur = coop_all_partials(V.part_b, n_part, sh);
  }
  // publish the scalars for the host check and the next launch: rz of iteration it0 + n_it and of the one before, r.r
  const double rr = coop_all_partials(part_c, n_part, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int it_end = it0 + n_it;
    __stcg(V.sc + (it_end & 1), rz_cur);
    __stcg(V.sc + ((it_end & 1) ^ 1), rz_prev);
    __stcg(V.sc + 2, rr);
    if (broken) __stcg(V.sc + 4, 1.0);
  }
}

inline int pcg_grid(const lrbms_context* ctx, int64_t n) {
  const int64_t want = (n + kPcgThreads / kLanesPerRow - 1) / (kPcgThreads / kLanesPerRow);
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->sm_count * 4));
}

inline size_t pcg_ws_doubles(int64_t n, int grid) { return (size_t)6 * n + (size_t)4 * grid + 8; }   // r z p Ap dinv p_alt

}  // namespace

extern "C" {

int lrbms_pcg_workspace_bytes(lrbms_handle_t h, int64_t n, size_t* bytes) {
  LRBMS_REQUIRE(h, h && bytes && n >= 0, "pcg_workspace_bytes: bad argument");
  *bytes = sizeof(double) * pcg_ws_doubles(n, pcg_grid(h, n));
  return LRBMS_OK;
}

int lrbms_pcg_solve(lrbms_handle_t h, int32_t n, const int32_t* rowptr, const int32_t* colind, const double* values,
                    const double* b, double* x, double rtol, int32_t max_iter, int32_t* iters_out, double* relres_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  LRBMS_REQUIRE(h, h && rowptr && colind && values && b && x && workspace, "pcg_solve: null argument");
  LRBMS_REQUIRE(h, n >= 0 && max_iter >= 0 && rtol >= 0.0, "pcg_solve: bad argument");
  const int grid = pcg_grid(h, n);
  LRBMS_REQUIRE(h, workspace_bytes >= sizeof(double) * pcg_ws_doubles(n, grid), "pcg_solve: workspace too small (see lrbms_pcg_workspace_bytes)");
  LRBMS_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (iters_out) *iters_out = 0;
  if (relres_out) *relres_out = 0.0;
  if (n == 0) return LRBMS_OK;
  double* w = (double*)workspace;
  PcgVecs V;
  V.r = w; V.z = w + n; V.p = w + 2 * (size_t)n; V.Ap = w + 3 * (size_t)n; V.dinv = w + 4 * (size_t)n;
  double* p_alt = w + 5 * (size_t)n;                       // second direction buffer of the cooperative kernel
  V.part_a = w + 6 * (size_t)n; V.part_b = V.part_a + grid;
  double* part_c = V.part_b + grid;
  double* part_d = part_c + grid;
  V.sc = part_d + grid;
  // cooperative path: one CTA of 1024 threads per SM at most (the whole grid must be resident for the grid-wide barriers)
  int coop_attr = 0, per_sm = 0;
  const bool cooperative = !h->single_launch_pcg_off &&
                           cudaDeviceGetAttribute(&coop_attr, cudaDevAttrCooperativeLaunch, h->device) == cudaSuccess && coop_attr &&
                           cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_iterate_kernel, kCoopThreads, 0) == cudaSuccess &&
                           per_sm >= 1;
  const int coop_grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)n * kLanesPerRow + kCoopThreads - 1) / kCoopThreads,
                                                                     std::min<int64_t>(h->sm_count, grid)));
  pcg_init_kernel<<<grid, kPcgThreads, 0, s>>>(n, rowptr, colind, values, b, x, V, part_d);
  pcg_init_scalars_kernel<<<1, kPcgThreads, 0, s>>>(grid, V, part_d);
  double sc[5] = {0, 0, 0, 0, 0};
  int it = 0;
  double relres = 0.0;
  for (;;) {
    LRBMS_CUDA_CHECK(h, cudaMemcpyAsync(sc, V.sc, sizeof(sc), cudaMemcpyDeviceToHost, s));
    LRBMS_CUDA_CHECK(h, cudaStreamSynchronize(s));
    relres = (sc[3] > 0.0) ? std::sqrt(sc[2] / sc[3]) : std::sqrt(sc[2]);
    if (sc[4] != 0.0) {
      if (iters_out) *iters_out = it;
      if (relres_out) *relres_out = relres;
      return lrbms_fail(h, LRBMS_ERR_NOT_SPD, "pcg_solve: p^T A p <= 0, the neighbourhood operator is not positive definite");
    }
    if (relres <= rtol || it >= max_iter) break;
    int n_it = std::min(kPcgCheck, max_iter - it);
    if (cooperative) {
      int it0 = it;
      void* args[] = {(void*)&n, (void*)&rowptr, (void*)&colind, (void*)&values, (void*)&x, (void*)&V, (void*)&p_alt,
                      (void*)&part_c, (void*)&it0, (void*)&n_it};
      LRBMS_CUDA_CHECK(h, cudaLaunchCooperativeKernel((const void*)pcg_iterate_kernel, dim3(coop_grid), dim3(kCoopThreads), args, 0, s));
      it += n_it;
      continue;
    }
    for (int k = 0; k < n_it; ++k, ++it) {
      const int parity = it & 1;
      pcg_spmv_kernel<<<grid, kPcgThreads, 0, s>>>(n, rowptr, colind, values, V);
      pcg_update_xr_kernel<<<grid, kPcgThreads, 0, s>>>(n, grid, parity, x, V, part_c);
      pcg_update_p_kernel<<<grid, kPcgThreads, 0, s>>>(n, grid, parity, V, part_c);
    }
  }
  LRBMS_CUDA_CHECK(h, cudaGetLastError());
  if (iters_out) *iters_out = it;
  if (relres_out) *relres_out = relres;
  return LRBMS_OK;
}

}  // extern "C"
