// Large reduced systems: mu-batched block-banded Cholesky out of HBM ("band" solver of the online plan).
//
// replaces: the dense numpy.linalg.solve behind rd.solve(mu) (reference online_enrichment.py:72,
// scripts/online_adaptive_lrbms.py:141) for reduced systems whose live factor window does not fit the shared memory of
// one SM (BASELINE configs[2]: 16x16 subdomains, n_red = 5 120; configs[3]: 8x8x8 subdomains, N = 40, n_red = 20 480,
// half bandwidth 2 600).  The shared-memory kernel (online.cu, solve_kernel_v2) gives one SM to one parameter; here one
// parameter's factor (hundreds of MB at configs[3]) lives in HBM and the whole GPU works on a chunk of parameters at once.
//
// Storage: the lower band of A(mu) / L in 64 x 64 blocks, block column J holding rows J .. J + kb.  A block is stored in
// DMMA *operand-fragment order* ("F-layout"): element (r, c) at ((c / 4) * 8 + r / 8) * 32 + (r % 8) * 4 + c % 4.  A 16-column
// slice of a block is 8 KB contiguous, so the update kernel stages its operands with one cp.async.bulk (TMA engine) per
// operand and k-chunk, completing on an mbarrier, and every warp-level fragment load is one conflict-free 256-byte
// shared-memory wavefront pair.  The same layout serves as the A operand (L_IK) and as the B operand (L_JK) of
// C -= L_IK L_JK^T.
//
// Left-looking, three launches per block column J over (targets, parameters of the chunk):
//   band_update_kernel  T_IJ = sum_q theta_q A_q[I,J] - sum_{K} L_IK L_JK^T         (FP64 tensor pipe, the bulk of the flops)
//   band_potrf_kernel   L_JJ = chol(T_JJ), W_J = L_JJ^{-1}                            (one CTA per parameter)
//   band_trsm_kernel    L_IJ = T_IJ W_J^T                                             (FP64 tensor pipe)
// then one launch of band_substitute_kernel (one CTA per parameter): y = L^{-1} f(mu), u = L^{-T} y.
#include <algorithm>
#include <cmath>

#include "band.h"

namespace {

constexpr int NB = kBandNB;                 // 64
constexpr int kBlk = NB * NB;               // doubles per block
constexpr int kUpdThreads = 256;
constexpr int kUpdStages = 4;
constexpr int kChunkDoubles = 16 * NB;      // one 16-column slice of a block in F-layout: 1024 doubles = 8 KB

__host__ __device__ __forceinline__ int f_off(int r, int c) { return (((c >> 2) * 8 + (r >> 3)) << 5) + ((r & 7) << 2) + (c & 3); }

struct BandParams {
  int32_t nbc, kb, n_red, n_pad, Q, Qf, n_theta, n_a;
  int64_t per_mu;            // doubles of workspace per parameter: band blocks, then the inverse diagonal blocks
  int64_t winv_off;          // offset of the inverse diagonal blocks inside a parameter's workspace
  const int32_t* a_map;      // [nbc * (kb + 1)]: compact operator block of band block (J, d), or -1
  const double* a_blocks;    // [Q][n_a][4096], F-layout; diagonal blocks hold both triangles
  const double* rhs;         // [Qf][n_pad]
};

__device__ __forceinline__ double* band_block(double* base, const BandParams& P, int J, int d) {
  return base + ((int64_t)J * (P.kb + 1) + d) * kBlk;
}

// ------------------------------------------------------------------------------------------------------
//  T_IJ = A_IJ(mu) - sum_{K = max(0, I - kb)}^{J - 1} L_IK L_JK^T          grid: (kb + 1 targets, parameters)
//  8 warps as 4 (rows) x 2 (columns): a warp owns 16 x 32 of the target = 2 x 4 DMMA tiles.
// ------------------------------------------------------------------------------------------------------
// 112 registers: two CTAs of this kernel leave register room for a CTA of band_potrf_kernel on the same SM, so that the
// look-ahead factorisation of the diagonal block really runs beside the off-diagonal updates
__global__ void __maxnreg__(112)
band_update_kernel(BandParams P, int J, int d0, const double* __restrict__ theta, double* __restrict__ work) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long full_bar[kUpdStages];
  const int d = d0 + blockIdx.x, I = J + d;
  if (I >= P.nbc) return;
  const int64_t mu = blockIdx.y;
  double* base = work + mu * P.per_mu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp >> 1, wc = warp & 1;
  const int K0 = max(0, I - P.kb);
  const int n_chunks = (J - K0) * 4;

  if (threadIdx.x == 0)
    for (int s = 0; s < kUpdStages; ++s) mbar_init(&full_bar[s], 1);
  fence_proxy_async();
  __syncthreads();

  auto issue = [&](int c) {
    const int K = K0 + (c >> 2), kc = c & 3, s = c % kUpdStages;
    double* dst = smem + s * 2 * kChunkDoubles;
    mbar_expect_tx(&full_bar[s], 2 * kChunkDoubles * 8);
    bulk_copy_g2s(dst, band_block(base, P, K, I - K) + kc * kChunkDoubles, kChunkDoubles * 8, &full_bar[s]);
    bulk_copy_g2s(dst + kChunkDoubles, band_block(base, P, K, J - K) + kc * kChunkDoubles, kChunkDoubles * 8, &full_bar[s]);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < kUpdStages - 1 && c < n_chunks; ++c) issue(c);

  // accumulators start from the assembled operator block: sum_q theta_q A_q, left to right (LincombOperator.assemble)
  double acc[2][4][2];
  {
    const int ai = P.a_map[J * (P.kb + 1) + d];
    const double* th = theta + mu * P.n_theta;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double v0 = 0.0, v1 = 0.0;
        if (ai >= 0) {
          const int r = 8 * (2 * wr + i) + g, c = 8 * (4 * wc + j) + 2 * t;
          const double* src = P.a_blocks + (int64_t)ai * kBlk + f_off(r, c);
          for (int q = 0; q < P.Q; ++q) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(src + (int64_t)q * P.n_a * kBlk));
            const double thq = __ldg(th + q);
            if (q == 0) { v0 = thq * v.x; v1 = thq * v.y; }
            else { v0 += thq * v.x; v1 += thq * v.y; }
          }
        }
        acc[i][j][0] = v0;
        acc[i][j][1] = v1;
      }
  }

  for (int c = 0; c < n_chunks; ++c) {
    // the slot of chunk c - 1 is free: everybody passed the barrier that closed iteration c - 1
    if (threadIdx.x == 0 && c + kUpdStages - 1 < n_chunks) issue(c + kUpdStages - 1);
    const int s = c % kUpdStages;
    mbar_wait(&full_bar[s], (unsigned)((c / kUpdStages) & 1));
    const double* As = smem + s * 2 * kChunkDoubles;
    const double* Bs = As + kChunkDoubles;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = -As[((ks * 8 + 2 * wr + i) << 5) + lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[((ks * 8 + 4 * wc + j) << 5) + lane];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }

  double* out = band_block(base, P, J, d);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 8 * (2 * wr + i) + g, c = 8 * (4 * wc + j) + 2 * t;
      *reinterpret_cast<double2*>(out + f_off(r, c)) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// ------------------------------------------------------------------------------------------------------
//  L_JJ = chol(T_JJ), W_J = L_JJ^{-1} (F-layout, the B operand of the triangular solves)     grid: parameters
// ------------------------------------------------------------------------------------------------------
constexpr int kPotrfLd = NB + 1;
__global__ void __launch_bounds__(256)
band_potrf_kernel(BandParams P, int J, double* __restrict__ work, int32_t* __restrict__ info) {
  extern __shared__ __align__(128) double smem[];
  double* S = smem;                       // the diagonal block, row-major; the factor overwrites its lower triangle and the
                                          // inverse W = L^-1 goes, transposed, into the (unused) strictly upper triangle:
                                          // W[i][j] (i > j) at S[j][i] -- one 33 KB buffer, so that a CTA of this kernel
                                          // fits beside two CTAs of band_update_kernel on one SM (look-ahead)
  __shared__ double wd[NB];               // diagonal of W
  __shared__ int s_bad;
  const int64_t mu = blockIdx.x;
  double* base = work + mu * P.per_mu;
  const double* T = band_block(base, P, J, 0);
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < kBlk; e += 256) {
    // F-layout element e = ((ks * 8 + rt) * 32 + g * 4 + t)
    const int t = e & 3, g = (e >> 2) & 7, rt = (e >> 5) & 7, ks = e >> 8;
    S[(8 * rt + g) * kPotrfLd + 4 * ks + t] = T[e];
  }
  __syncthreads();
  const int row = tid & 63, part = tid >> 6;
  for (int k = 0; k < NB; ++k) {
    double akk = S[k * kPotrfLd + k];
    if (NB * J + k >= P.n_red) akk = 1.0;                     // padding rows: identity
    if (!(akk > 0.0)) { if (tid == 0 && s_bad == 0) s_bad = NB * J + k + 1; akk = 1.0; }
    const double rinv = rsqrt(akk);
    double lik = 0.0;
    if (part == 0) {
      if (row > k) lik = S[row * kPotrfLd + k] * rinv;
      else if (row == k) lik = akk * rinv;
    }
    __syncthreads();                                          // everybody has read the pivot
    if (part == 0 && row >= k) S[row * kPotrfLd + k] = lik;
    __syncthreads();
    if (row > k) {
      const double li = S[row * kPotrfLd + k];
      for (int j = k + 1 + part; j <= row; j += 4) S[row * kPotrfLd + j] -= li * S[j * kPotrfLd + k];
    }
    __syncthreads();
  }
  // inverse: four threads per column j of W, forward substitution  W[i][j] = (delta_ij - sum_{k<i} L[i][k] W[k][j]) / L[i][i]
  {
    const int j = tid >> 2, p4 = tid & 3;
    double* wj = S + j * kPotrfLd;                            // W[k][j], k > j, at wj[k] (upper triangle, row j)
    if (p4 == 0) wd[j] = 1.0 / S[j * kPotrfLd + j];
    __syncwarp();
    for (int i = 1; i < NB; ++i) {                            // uniform trip count: the shuffles are warp-wide
      double s = 0.0;
      for (int k = j + p4; k < i; k += 4) s += S[i * kPotrfLd + k] * ((k == j) ? wd[j] : wj[k]);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (p4 == 0 && i > j) wj[i] = -s / S[i * kPotrfLd + i];
      __syncwarp();
    }
  }
  __syncthreads();
  double* W = base + P.winv_off + (int64_t)J * kBlk;
  for (int e = tid; e < kBlk; e += 256) {
    const int t = e & 3, g = (e >> 2) & 7, rt = (e >> 5) & 7, ks = e >> 8;
    const int r = 8 * rt + g, c = 4 * ks + t;                 // W[r][c]
    W[e] = (r == c) ? wd[r] : (r > c ? S[c * kPotrfLd + r] : 0.0);
  }
  if (tid == 0 && s_bad && info && info[mu] == 0) info[mu] = s_bad;
}

// ------------------------------------------------------------------------------------------------------
//  L_IJ = T_IJ W_J^T  (in place)                                             grid: (kb targets, parameters)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kUpdThreads, 2)
band_trsm_kernel(BandParams P, int J, double* __restrict__ work) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long bar;
  const int d = blockIdx.x + 1, I = J + d;
  if (I >= P.nbc) return;
  const int64_t mu = blockIdx.y;
  double* base = work + mu * P.per_mu;
  double* TL = band_block(base, P, J, d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp >> 1, wc = warp & 1;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 2 * kBlk * 8);
    bulk_copy_g2s(smem, TL, kBlk * 8, &bar);
    bulk_copy_g2s(smem + kBlk, base + P.winv_off + (int64_t)J * kBlk, kBlk * 8, &bar);
  }
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  mbar_wait(&bar, 0);
  const double* As = smem;
  const double* Bs = smem + kBlk;
  // W is lower triangular: W[c][k] = 0 for k > c, so output columns 32 wc .. 32 wc + 31 need k < 32 (wc + 1) only
  const int ks_end = 8 * (wc + 1);
  for (int ks = 0; ks < ks_end; ++ks) {
    double a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = As[((ks * 8 + 2 * wr + i) << 5) + lane];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = Bs[((ks * 8 + 4 * wc + j) << 5) + lane];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 8 * (2 * wr + i) + g, c = 8 * (4 * wc + j) + 2 * t;
      *reinterpret_cast<double2*>(TL + f_off(r, c)) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// ------------------------------------------------------------------------------------------------------
//  y = L^{-1} f(mu),  u = L^{-T} y.   One CTA per parameter; y / u live in the parameter's row of the output array.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
band_substitute_kernel(BandParams P, const double* __restrict__ theta, double* __restrict__ work, double* __restrict__ u) {
  __shared__ double sx[(kBandMaxKb + 1) * NB];   // ring of the last kb + 1 solution blocks (block B lives in slot B % (kb + 1))
  __shared__ double part[16 * NB];
  __shared__ double tv[NB];
  const int64_t mu = blockIdx.x;
  double* base = work + mu * P.per_mu;
  double* um = u + mu * P.n_red;
  const double* th = theta + mu * P.n_theta + P.Q;
  const int tid = threadIdx.x;
  const int ring = P.kb + 1;
  // ---- forward: y_J = W_J (f_J - sum_{K = J - kb}^{J - 1} L_JK y_K);  thread (r, kq): row r, k-steps kq, kq + 4, ...
  {
    const int r = tid & 63, kq = tid >> 6;
    for (int J = 0; J < P.nbc; ++J) {
      double s = 0.0;
      for (int K = max(0, J - P.kb); K < J; ++K) {
        const double* Lb = band_block(base, P, K, J - K);
        const double* y = sx + (K % ring) * NB;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ks = kq + 4 * i;
          const double4 l = *reinterpret_cast<const double4*>(Lb + ((ks * 8 + (r >> 3)) << 5) + ((r & 7) << 2));
          s += l.x * y[4 * ks] + l.y * y[4 * ks + 1] + l.z * y[4 * ks + 2] + l.w * y[4 * ks + 3];
        }
      }
      part[kq * NB + r] = s;
      __syncthreads();
      if (tid < NB) {
        double f = 0.0;
        for (int q = 0; q < P.Qf; ++q) f += th[q] * P.rhs[(int64_t)q * P.n_pad + NB * J + tid];
        tv[tid] = f - ((part[tid] + part[NB + tid]) + (part[2 * NB + tid] + part[3 * NB + tid]));
      }
      __syncthreads();
      {
        const double* Wb = base + P.winv_off + (int64_t)J * kBlk;
        double w = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ks = kq + 4 * i;
          const double4 l = *reinterpret_cast<const double4*>(Wb + ((ks * 8 + (r >> 3)) << 5) + ((r & 7) << 2));
          w += l.x * tv[4 * ks] + l.y * tv[4 * ks + 1] + l.z * tv[4 * ks + 2] + l.w * tv[4 * ks + 3];
        }
        part[kq * NB + r] = w;
      }
      __syncthreads();
      if (tid < NB) {
        const double yv = (part[tid] + part[NB + tid]) + (part[2 * NB + tid] + part[3 * NB + tid]);
        sx[(J % ring) * NB + tid] = yv;
        if (NB * J + tid < P.n_red) um[NB * J + tid] = yv;        // y parked in the output row
      }
      __syncthreads();
    }
  }
  // ---- backward: u_J = W_J^T (y_J - sum_{I = J + 1}^{J + kb} L_IJ^T u_I);  thread (ks, rt, half): 4 columns x 4 rows of a block
  {
    const int ks = tid >> 4, rt = (tid >> 1) & 7, half = tid & 1;
    for (int J = P.nbc - 1; J >= 0; --J) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int I = J + 1; I <= min(J + P.kb, P.nbc - 1); ++I) {
        const double* Lb = band_block(base, P, J, I - J) + ((ks * 8 + rt) << 5) + half * 16;
        const double* x = sx + (I % ring) * NB + 8 * rt + 4 * half;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const double4 l = *reinterpret_cast<const double4*>(Lb + 4 * gg);
          const double xv = x[gg];
          s[0] += l.x * xv; s[1] += l.y * xv; s[2] += l.z * xv; s[3] += l.w * xv;
        }
      }
      // the ring slot of block J still holds y_J (block J + kb + 1 used the same slot and is no longer needed)
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 4; ++c) part[(rt * 2 + half) * NB + 4 * ks + c] = s[c];
      __syncthreads();
      if (tid < NB) {
        double v = 0.0;
        for (int p = 0; p < 16; ++p) v += part[p * NB + tid];
        const double yv = (NB * J + tid < P.n_red) ? um[NB * J + tid] : 0.0;
        tv[tid] = yv - v;
      }
      __syncthreads();
      {
        const double* Wb = base + P.winv_off + (int64_t)J * kBlk + ((ks * 8 + rt) << 5) + half * 16;
        double w[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const double4 l = *reinterpret_cast<const double4*>(Wb + 4 * gg);
          const double xv = tv[8 * rt + 4 * half + gg];
          w[0] += l.x * xv; w[1] += l.y * xv; w[2] += l.z * xv; w[3] += l.w * xv;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) part[(rt * 2 + half) * NB + 4 * ks + c] = w[c];
      }
      __syncthreads();
      if (tid < NB) {
        double v = 0.0;
        for (int p = 0; p < 16; ++p) v += part[p * NB + tid];
        sx[(J % ring) * NB + tid] = v;
        if (NB * J + tid < P.n_red) um[NB * J + tid] = v;
      }
      __syncthreads();
    }
  }
}

__global__ void band_clear_info_kernel(int64_t n, int32_t* info) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) info[i] = 0;
}

BandParams to_params(const lrbms_band_plan& B) {
  BandParams P;
  P.nbc = B.nbc; P.kb = B.kb; P.n_red = B.n_red; P.n_pad = B.n_pad; P.Q = B.Q; P.Qf = B.Qf; P.n_theta = B.Q + B.Qf; P.n_a = B.n_a;
  P.per_mu = B.per_mu_doubles; P.winv_off = B.winv_off; P.a_map = B.d_a_map; P.a_blocks = B.d_a_blocks; P.rhs = B.d_rhs;
  return P;
}

}  // namespace

int lrbms_band_build(lrbms_plan* plan, lrbms_band_plan& B, int32_t n_sub, const int32_t* sizes, const int32_t* offsets, int32_t Q,
                     int32_t Qf, int32_t n_blocks, const int32_t* bi, const int32_t* bj, const int64_t* block_offset,
                     const double* host_blocks, const double* host_rhs) {
  lrbms_context* ctx = plan->ctx;
  const int n_red = offsets[n_sub];
  B.n_red = n_red;
  B.n_pad = (n_red + NB - 1) / NB * NB;
  B.nbc = B.n_pad / NB;
  B.Q = Q;
  B.Qf = Qf;
  // scalar half bandwidth of the stored pattern -> block half bandwidth
  int hb = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const int i = bi[b], j = bj[b];
    if (i < j || sizes[i] == 0 || sizes[j] == 0) continue;
    hb = std::max(hb, offsets[i + 1] - 1 - offsets[j]);
  }
  B.half_bandwidth = hb;
  B.kb = std::min(B.nbc - 1, (hb + NB - 1) / NB);
  if (B.kb > kBandMaxKb) return lrbms_fail(ctx, LRBMS_ERR_UNSUPPORTED, "online_plan_create: band wider than the band solver supports");
  const int kb1 = B.kb + 1;
  B.winv_off = (int64_t)B.nbc * kb1 * kBlk;
  B.per_mu_doubles = B.winv_off + (int64_t)B.nbc * kBlk;
  // ---- compact operator blocks in F-layout
  std::vector<int32_t> a_map((size_t)B.nbc * kb1, -1);
  int n_a = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const int i = bi[b], j = bj[b];
    if (i < j || sizes[i] == 0 || sizes[j] == 0) continue;
    for (int I = offsets[i] / NB; I <= (offsets[i + 1] - 1) / NB; ++I)
      for (int J = offsets[j] / NB; J <= std::min(I, (offsets[j + 1] - 1) / NB); ++J) {
        int32_t& slot = a_map[(size_t)J * kb1 + (I - J)];
        if (slot < 0) slot = n_a++;
      }
  }
  B.n_a = std::max(1, n_a);
  std::vector<double> blocks((size_t)Q * B.n_a * kBlk, 0.0);
  for (int q = 0; q < Q; ++q)
    for (int b = 0; b < n_blocks; ++b) {
      const int i = bi[b], j = bj[b];
      if (i < j) continue;
      const double* blk = host_blocks + block_offset[(int64_t)q * n_blocks + b];
      const int Ni = sizes[i], Nj = sizes[j];
      for (int a = 0; a < Ni; ++a)
        for (int c = 0; c < Nj; ++c) {
          const int r = offsets[i] + a, cc = offsets[j] + c;
          if (r < cc) continue;
          const int I = r / NB, J = cc / NB;
          double* dst = blocks.data() + ((size_t)q * B.n_a + a_map[(size_t)J * kb1 + (I - J)]) * kBlk;
          dst[f_off(r % NB, cc % NB)] = blk[(int64_t)a * Nj + c];
          if (I == J) dst[f_off(cc % NB, r % NB)] = blk[(int64_t)a * Nj + c];
        }
    }
  std::vector<double> rhs((size_t)Qf * B.n_pad, 0.0);
  for (int q = 0; q < Qf; ++q) std::copy(host_rhs + (size_t)q * n_red, host_rhs + (size_t)(q + 1) * n_red, rhs.begin() + (size_t)q * B.n_pad);
  int32_t* d_i32 = nullptr;
  double* d_f64 = nullptr;
  int rc = plan_upload(plan, &d_i32, a_map);
  if (rc) return rc;
  B.d_a_map = d_i32;
  rc = plan_upload(plan, &d_f64, blocks);
  if (rc) return rc;
  B.d_a_blocks = d_f64;
  rc = plan_upload(plan, &d_f64, rhs);
  if (rc) return rc;
  B.d_rhs = d_f64;
  // executed flops per parameter (the band is treated as dense): updates + triangular solves + diagonal factorisations
  double fl = 0;
  for (int J = 0; J < B.nbc; ++J)
    for (int d = 0; d <= B.kb && J + d < B.nbc; ++d) {
      fl += 2.0 * kBlk * NB * (J - std::max(0, J + d - B.kb));
      fl += (d == 0) ? (double)NB * NB * NB / 3.0 : (double)kBlk * NB;
    }
  B.flops_per_mu = fl;
  B.upd_smem = sizeof(double) * kUpdStages * 2 * kChunkDoubles;
  B.trsm_smem = sizeof(double) * 2 * kBlk;
  B.potrf_smem = sizeof(double) * NB * kPotrfLd;
  cudaError_t e = cudaFuncSetAttribute(band_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.upd_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(band_potrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.potrf_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(band_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.trsm_smem);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&B.side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&B.ev_diag, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&B.ev_potrf, cudaEventDisableTiming);
  if (e != cudaSuccess) return lrbms_fail(ctx, LRBMS_ERR_CUDA, std::string("band solver: ") + cudaGetErrorString(e));
  return LRBMS_OK;
}

void lrbms_band_release(lrbms_band_plan& B) {
  if (B.ev_diag) cudaEventDestroy(B.ev_diag);
  if (B.ev_potrf) cudaEventDestroy(B.ev_potrf);
  if (B.side) cudaStreamDestroy(B.side);
  B.ev_diag = B.ev_potrf = nullptr;
  B.side = nullptr;
}

int64_t lrbms_band_chunk(const lrbms_band_plan& B, int64_t n_mu, size_t workspace_bytes) {
  const int64_t fit = (int64_t)(workspace_bytes / (sizeof(double) * (size_t)B.per_mu_doubles));
  return std::min<int64_t>(std::min<int64_t>(n_mu, fit), 65535);
}

size_t lrbms_band_workspace_bytes(const lrbms_band_plan& B, int64_t n_mu) {
  // the whole batch if it fits the budget, else as many parameters per chunk as the budget allows (at least one)
  const size_t per = sizeof(double) * (size_t)B.per_mu_doubles;
  int64_t chunk = std::max<int64_t>(1, (int64_t)(kBandWorkspaceBudget / per));
  chunk = std::min<int64_t>(std::min<int64_t>(chunk, std::max<int64_t>(1, n_mu)), 65535);
  return per * (size_t)chunk;
}

int lrbms_band_solve(lrbms_context* ctx, const lrbms_band_plan& B, int64_t n_mu, const double* theta, double* u, int32_t* info,
                     void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const int64_t max_chunk = lrbms_band_chunk(B, n_mu, workspace_bytes);
  if (max_chunk < 1) return lrbms_fail(ctx, LRBMS_ERR_INVALID, "online_solve: workspace too small (see lrbms_online_workspace_bytes)");
  // equal chunks (64 parameters with room for 57 factors: 32 + 32, not 57 + 7 -- a small tail chunk leaves most SMs idle)
  const int64_t n_chunks = (n_mu + max_chunk - 1) / max_chunk;
  const int64_t chunk = (n_mu + n_chunks - 1) / n_chunks;
  BandParams P = to_params(B);
  double* work = reinterpret_cast<double*>(workspace);
  const int n_theta = B.Q + B.Qf;
  if (info) band_clear_info_kernel<<<(unsigned)((n_mu + 255) / 256), 256, 0, s>>>(n_mu, info);
  for (int64_t lo = 0; lo < n_mu; lo += chunk) {
    const int64_t m = std::min<int64_t>(chunk, n_mu - lo);
    const double* th = theta + lo * n_theta;
    for (int J = 0; J < B.nbc; ++J) {
      const int n_t = std::min(B.kb, B.nbc - 1 - J) + 1;
      // look-ahead: the diagonal target and its factorisation (one latency-bound CTA per parameter) run on the side stream
      // beside the off-diagonal updates of the column, which do not need them; the triangular solves wait for both
      cudaEventRecord(B.ev_diag, s);                       // everything up to the triangular solves of column J - 1
      cudaStreamWaitEvent(B.side, B.ev_diag, 0);
      band_update_kernel<<<dim3(1, (unsigned)m), kUpdThreads, B.upd_smem, B.side>>>(P, J, 0, th, work);
      band_potrf_kernel<<<(unsigned)m, 256, B.potrf_smem, B.side>>>(P, J, work, info ? info + lo : nullptr);
      cudaEventRecord(B.ev_potrf, B.side);
      if (n_t > 1) band_update_kernel<<<dim3((unsigned)(n_t - 1), (unsigned)m), kUpdThreads, B.upd_smem, s>>>(P, J, 1, th, work);
      cudaStreamWaitEvent(s, B.ev_potrf, 0);
      if (n_t > 1) band_trsm_kernel<<<dim3((unsigned)(n_t - 1), (unsigned)m), kUpdThreads, B.trsm_smem, s>>>(P, J, work);
    }
    band_substitute_kernel<<<(unsigned)m, 256, 0, s>>>(P, th, work, u + lo * B.n_red);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return lrbms_fail(ctx, LRBMS_ERR_CUDA, std::string("band solver: ") + cudaGetErrorString(e));
  return LRBMS_OK;
}
