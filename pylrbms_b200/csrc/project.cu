// Offline half of the LRBMS hot path: batched block-CSR SpMM (K1) and the fused Galerkin projection
// G = alpha * VL^T (A VR) (K1 + K2) on FP64 tensor cores (DMMA.8x8x4).
//
// Fragment-direct design.  One warp owns a k-step of 4 consecutive rows.  Lane 4*g + t (g = 0..7, t = 0..3)
//   * computes (A VR)[k0 + t][8n + g] for n < NT directly in registers: the CSR row is walked by the eight lanes
//     that share t (broadcast loads of value / column index), each gathering one 64-byte segment of the
//     dof-major VR row per column tile  -> that *is* the DMMA B fragment, no shared-memory staging, no barrier;
//   * loads VL[k0 + t][8m + g] for m < MT                                        -> the DMMA A fragment;
//   * issues MT x NT DMMA.8x8x4, accumulating the (8 MT) x (8 NT) piece of G in registers over all its rows.
// A CTA (8 warps) covers a contiguous row range; warps are reduced through shared memory in fixed order, CTAs of
// the same output through a global scratch buffer summed by the last CTA to arrive, again in fixed order, so the
// result is bit-reproducible run to run.  Row groups of 4 with no non-zeros (coupling blocks: only interface rows
// are populated) are skipped without touching VL.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxTile = 5;                               // largest MT / NT instantiated: 40 x 40 output chunk per CTA
constexpr int64_t kPartialStride = kMaxTile * kMaxTile * 64;  // doubles per partial slot (same for every launch)

struct ProjItem {
  int32_t desc;       // descriptor index
  int32_t l0, c0;     // first column of the VL / VR chunk
  int32_t row0, row1; // row range (multiple of 4 except at the end)
  int32_t group;      // output group (desc, l-chunk, r-chunk)
  int32_t slot;       // index of this item inside its group
  int32_t group_size;
  int32_t mirror;     // symmetric descriptor, off-diagonal chunk pair: also write the transposed chunk
};

struct DevDesc {       // device-side mirror of lrbms_project_desc_t (VR/rowptr possibly redirected to scratch)
  const int32_t* rowptr;
  const int32_t* colind;
  const double* values;
  int32_t n_rows, n_cols;
  const double* VL; int32_t ldl, NL;
  const double* VR; int32_t ldr, NR;
  double* out; int32_t ldo;
  double alpha;
};

template <int MT, int NT, bool HAS_A>
__global__ void __launch_bounds__(kThreads, (MT * NT <= 9) ? 3 : ((MT * NT <= 16) ? 2 : 1))
project_kernel(const ProjItem* __restrict__ items, const DevDesc* __restrict__ descs, double* __restrict__ partials,
               int32_t* __restrict__ flags, int32_t* __restrict__ counters, const int64_t* __restrict__ group_partial_base) {
  const ProjItem it = items[blockIdx.x];
  const DevDesc D = descs[it.desc];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nl = min(8 * MT, D.NL - it.l0), nr = min(8 * NT, D.NR - it.c0);

  double acc[MT][NT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;

  bool any_work = false;
  const double* __restrict__ VL = D.VL + it.l0 + g;
  const double* __restrict__ VR = D.VR + it.c0 + g;

  if (HAS_A) {
    for (int k0 = it.row0 + 4 * warp; k0 < it.row1; k0 += 4 * kWarps) {
      const int row = k0 + t;
      const bool row_ok = row < it.row1;
      double b[NT];
#pragma unroll
      for (int n = 0; n < NT; ++n) b[n] = 0.0;
      int p0 = 0, len = 0;
      if (row_ok) {
        p0 = D.rowptr[row];
        len = D.rowptr[row + 1] - p0;
      }
      const int maxlen = __reduce_max_sync(0xffffffffu, len);
      if (maxlen == 0) continue;
      const int32_t* __restrict__ ci = D.colind + p0;
      const double* __restrict__ va = D.values + p0;
#pragma unroll 2
      for (int p = 0; p < maxlen; ++p) {
        if (p < len) {
          const double a = va[p];
          const double* __restrict__ vr = VR + (int64_t)ci[p] * D.ldr;
#pragma unroll
          for (int n = 0; n < NT; ++n)
            if (8 * n + g < nr) b[n] = fma(a, vr[8 * n], b[n]);
        }
      }
      any_work = true;
      double a[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) a[m] = 0.0;
      if (row_ok) {
        const double* __restrict__ vl = VL + (int64_t)row * D.ldl;
#pragma unroll
        for (int m = 0; m < MT; ++m)
          if (8 * m + g < nl) a[m] = vl[8 * m];
      }
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma884(acc[m][n][0], acc[m][n][1], a[m], b[n]);
    }
  } else {
    // dense G = VL^T VR: register double buffering -- the fragments of the next k-step are in flight while the
    // MT x NT DMMAs of the current one issue
    auto load_frags = [&](int k0, double (&a)[MT], double (&b)[NT]) {
      const int row = k0 + t;
      const bool row_ok = row < it.row1;
      const double* __restrict__ vl = VL + (int64_t)row * D.ldl;
      const double* __restrict__ vr = VR + (int64_t)row * D.ldr;
#pragma unroll
      for (int m = 0; m < MT; ++m) a[m] = (row_ok && 8 * m + g < nl) ? vl[8 * m] : 0.0;
#pragma unroll
      for (int n = 0; n < NT; ++n) b[n] = (row_ok && 8 * n + g < nr) ? vr[8 * n] : 0.0;
    };
    int k0 = it.row0 + 4 * warp;
    if (k0 < it.row1) {
      any_work = true;
      double a_cur[MT], b_cur[NT], a_nxt[MT], b_nxt[NT];
      load_frags(k0, a_cur, b_cur);
      for (; k0 < it.row1; k0 += 4 * kWarps) {
        const int k1 = k0 + 4 * kWarps;
        if (k1 < it.row1) load_frags(k1, a_nxt, b_nxt);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int n = 0; n < NT; ++n) dmma884(acc[m][n][0], acc[m][n][1], a_cur[m], b_cur[n]);
#pragma unroll
        for (int m = 0; m < MT; ++m) a_cur[m] = a_nxt[m];
#pragma unroll
        for (int n = 0; n < NT; ++n) b_cur[n] = b_nxt[n];
      }
    }
  }

  // ---- reduce the 8 warps in fixed order through shared memory
  __shared__ double red[MT * NT * 64];
  __shared__ int s_any, s_last;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  if (any_work && lane == 0) atomicOr(&s_any, 1);
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          double* r = red + (m * NT + n) * 64 + g * 8 + 2 * t;
          if (w == 0) { r[0] = acc[m][n][0]; r[1] = acc[m][n][1]; }
          else { r[0] += acc[m][n][0]; r[1] += acc[m][n][1]; }
        }
    }
    __syncthreads();
  }
  const bool cta_any = s_any != 0;

  // ---- combine the CTAs of this output group
  if (it.group_size == 1) {
    for (int e = threadIdx.x; e < nl * nr; e += kThreads) {
      const int a = e / nr, bb = e - a * nr;
      const double v = D.alpha * red[((a >> 3) * NT + (bb >> 3)) * 64 + (a & 7) * 8 + (bb & 7)];
      D.out[(int64_t)(it.l0 + a) * D.ldo + it.c0 + bb] = v;
      if (it.mirror) D.out[(int64_t)(it.c0 + bb) * D.ldo + it.l0 + a] = v;
    }
    return;
  }
  const int64_t base = group_partial_base[it.group];
  double* my = partials + (base + it.slot) * kPartialStride;
  if (cta_any)
    for (int e = threadIdx.x; e < MT * NT * 64; e += kThreads) my[e] = red[e];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    flags[base + it.slot] = cta_any ? 1 : 0;
    __threadfence();
    const int prev = atomicAdd(&counters[it.group], 1);
    s_last = (prev == it.group_size - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int e = threadIdx.x; e < nl * nr; e += kThreads) {
    const int a = e / nr, bb = e - a * nr;
    const int idx = ((a >> 3) * NT + (bb >> 3)) * 64 + (a & 7) * 8 + (bb & 7);
    double s = 0.0;
    for (int sl = 0; sl < it.group_size; ++sl)
      if (__ldcg(&flags[base + sl])) s += __ldcg(&partials[(base + sl) * kPartialStride + idx]);
    D.out[(int64_t)(it.l0 + a) * D.ldo + it.c0 + bb] = D.alpha * s;
    if (it.mirror) D.out[(int64_t)(it.c0 + bb) * D.ldo + it.l0 + a] = D.alpha * s;
  }
  __syncthreads();
  if (threadIdx.x == 0) counters[it.group] = 0;   // self-cleaning for the next run
}

// ------------------------------------------------------------------------------------------------------
//  SpMM  W = A V  in the same (row t, column 8n + g) lane layout; one warp per 4 rows and 8*NT columns.
// ------------------------------------------------------------------------------------------------------
struct SpmmItem { int32_t desc, row0, row1, c0; };

template <int NT>
__global__ void __launch_bounds__(kThreads) spmm_kernel(const SpmmItem* __restrict__ items,
                                                        const lrbms_spmm_desc_t* __restrict__ descs) {
  const SpmmItem it = items[blockIdx.x];
  const lrbms_spmm_desc_t D = descs[it.desc];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nr = min(8 * NT, D.N - it.c0);
  const double* __restrict__ V = D.V + it.c0 + g;
  for (int k0 = it.row0 + 4 * warp; k0 < it.row1; k0 += 4 * kWarps) {
    const int row = k0 + t;
    if (row >= it.row1) continue;
    double b[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) b[n] = 0.0;
    const int p0 = D.rowptr[row], p1 = D.rowptr[row + 1];
#pragma unroll 2
    for (int p = p0; p < p1; ++p) {
      const double a = D.values[p];
      const double* __restrict__ vr = V + (int64_t)D.colind[p] * D.ldv;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (8 * n + g < nr) b[n] = fma(a, vr[8 * n], b[n]);
    }
    double* __restrict__ w = D.W + (int64_t)row * D.ldw + it.c0 + g;
#pragma unroll
    for (int n = 0; n < NT; ++n)
      if (8 * n + g < nr) w[8 * n] = b[n];
  }
}

// per-descriptor statistics for the roofline accounting (run once at plan creation)
__global__ void csr_stats_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, int n_rows,
                                 unsigned char* __restrict__ colflag, unsigned long long* __restrict__ out) {
  unsigned long long nonempty = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
    const int p0 = rowptr[r], p1 = rowptr[r + 1];
    if (p1 > p0) ++nonempty;
    for (int p = p0; p < p1; ++p) colflag[colind[p]] = 1;
  }
  atomicAdd(&out[0], nonempty);
}
__global__ void count_flags_kernel(const unsigned char* __restrict__ colflag, int n, unsigned long long* __restrict__ out) {
  unsigned long long c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += colflag[i];
  atomicAdd(&out[1], c);
}

// ------------------------------------------------------------------------------------------------------
//  plans
// ------------------------------------------------------------------------------------------------------
struct SpmmLaunch {
  int nt;
  SpmmItem* d_items = nullptr;
  int n_items = 0;
};

struct SpmmPlan : lrbms_plan {
  lrbms_spmm_desc_t* d_descs = nullptr;
  std::vector<SpmmLaunch> launches;
  int run(void* stream) override;
};

template <int NT>
static void launch_spmm(const SpmmLaunch& L, const lrbms_spmm_desc_t* d_descs, cudaStream_t s) {
  spmm_kernel<NT><<<L.n_items, kThreads, 0, s>>>(L.d_items, d_descs);
}

int SpmmPlan::run(void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  for (const SpmmLaunch& L : launches) {
    if (!L.n_items) continue;
    switch (L.nt) {
      case 1: launch_spmm<1>(L, d_descs, s); break;
      case 2: launch_spmm<2>(L, d_descs, s); break;
      case 3: launch_spmm<3>(L, d_descs, s); break;
      case 4: launch_spmm<4>(L, d_descs, s); break;
      case 5: launch_spmm<5>(L, d_descs, s); break;
      case 6: launch_spmm<6>(L, d_descs, s); break;
      case 7: launch_spmm<7>(L, d_descs, s); break;
      default: launch_spmm<8>(L, d_descs, s); break;
    }
  }
  LRBMS_CUDA_CHECK(ctx, cudaGetLastError());
  return LRBMS_OK;
}

// build the SpMM work list for a set of descriptors; shared by the SpMM plan and by the two-step projection
static void build_spmm_items(const std::vector<lrbms_spmm_desc_t>& descs, int sm_count,
                             std::vector<std::vector<SpmmItem>>& by_nt /* index nt-1 */) {
  by_nt.assign(8, {});
  // rows per CTA: aim at >= 4 CTAs per SM over the whole batch, in multiples of 32 rows (one k-step per warp)
  int64_t total_rows = 0;
  for (const auto& d : descs) total_rows += (int64_t)d.n_rows * ((d.N + 63) / 64);
  int64_t target_ctas = (int64_t)sm_count * 8;
  int64_t rows_per_cta = std::max<int64_t>(32, ((total_rows / std::max<int64_t>(1, target_ctas) + 31) / 32) * 32);
  rows_per_cta = std::min<int64_t>(rows_per_cta, 1024);
  for (size_t i = 0; i < descs.size(); ++i) {
    const auto& d = descs[i];
    for (int c0 = 0; c0 < d.N; c0 += 64) {
      const int nt = (std::min(64, d.N - c0) + 7) / 8;
      for (int64_t r0 = 0; r0 < d.n_rows; r0 += rows_per_cta) {
        SpmmItem it{(int32_t)i, (int32_t)r0, (int32_t)std::min<int64_t>(d.n_rows, r0 + rows_per_cta), c0};
        by_nt[nt - 1].push_back(it);
      }
    }
  }
}

struct ProjLaunch {
  int mt, nt;
  bool has_a;
  ProjItem* d_items = nullptr;
  int n_items = 0;
};

struct ProjectPlan : lrbms_plan {
  DevDesc* d_descs = nullptr;
  double* d_partials = nullptr;
  int32_t* d_flags = nullptr;
  int32_t* d_counters = nullptr;
  int64_t* d_group_base = nullptr;
  std::vector<ProjLaunch> launches;
  // two-step path (large N with a sparse operator): SpMM into plan-owned scratch first
  lrbms_spmm_desc_t* d_spmm_descs = nullptr;
  std::vector<SpmmLaunch> spmm_launches;
  std::vector<lrbms_project_desc_t> host_descs;
  bool accounted = false;
  int run(void* stream) override;
  void ensure_info() override;
};

template <int MT, int NT, bool HAS_A>
static void launch_project(const ProjLaunch& L, const ProjectPlan* P, cudaStream_t s) {
  project_kernel<MT, NT, HAS_A><<<L.n_items, kThreads, 0, s>>>(L.d_items, P->d_descs, P->d_partials, P->d_flags,
                                                                P->d_counters, P->d_group_base);
}

template <int MT, bool HAS_A>
static void dispatch_nt(const ProjLaunch& L, const ProjectPlan* P, cudaStream_t s) {
  switch (L.nt) {
    case 1: launch_project<MT, 1, HAS_A>(L, P, s); break;
    case 2: launch_project<MT, 2, HAS_A>(L, P, s); break;
    case 3: launch_project<MT, 3, HAS_A>(L, P, s); break;
    case 4: launch_project<MT, 4, HAS_A>(L, P, s); break;
    default: launch_project<MT, 5, HAS_A>(L, P, s); break;
  }
}

template <bool HAS_A>
static void dispatch_mt(const ProjLaunch& L, const ProjectPlan* P, cudaStream_t s) {
  switch (L.mt) {
    case 1: dispatch_nt<1, HAS_A>(L, P, s); break;
    case 2: dispatch_nt<2, HAS_A>(L, P, s); break;
    case 3: dispatch_nt<3, HAS_A>(L, P, s); break;
    case 4: dispatch_nt<4, HAS_A>(L, P, s); break;
    default: dispatch_nt<5, HAS_A>(L, P, s); break;
  }
}

// accounting (tight and SURVEY-formula byte counts) from the device CSR data; lazy because it costs two small kernels
// and two synchronous copies per descriptor
void ProjectPlan::ensure_info() {
  if (accounted) return;
  accounted = true;
  unsigned long long* d_cnt = nullptr;
  unsigned char* d_flag = nullptr;
  int max_cols = 1;
  for (const auto& d : host_descs) max_cols = std::max(max_cols, d.n_cols);
  if (cudaMalloc((void**)&d_cnt, 2 * sizeof(unsigned long long)) != cudaSuccess) return;
  if (cudaMalloc((void**)&d_flag, (size_t)max_cols) != cudaSuccess) { cudaFree(d_cnt); return; }
  for (const auto& d : host_descs) {
    double nl = d.NL, nr = d.NR, r = d.n_rows, c = d.n_cols;
    if (!d.rowptr) {
      double b = 8.0 * r * (nl + nr) + 8.0 * nl * nr;
      info_bytes += b; info_bytes_survey += b; info_flops += 2.0 * r * nl * nr;
      continue;
    }
    int32_t nnz = 0;
    unsigned long long cnt[2] = {0, 0};
    if (d.n_rows > 0) {
      cudaMemcpy(&nnz, d.rowptr + d.n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost);
      cudaMemset(d_cnt, 0, 2 * sizeof(unsigned long long));
      cudaMemset(d_flag, 0, (size_t)std::max(1, d.n_cols));
      csr_stats_kernel<<<64, 256>>>(d.rowptr, d.colind, d.n_rows, d_flag, d_cnt);
      count_flags_kernel<<<64, 256>>>(d_flag, d.n_cols, d_cnt);
      cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost);
    }
    info_bytes += 12.0 * nnz + 4.0 * (r + 1) + 8.0 * (double)cnt[1] * nr + 8.0 * (double)cnt[0] * nl + 8.0 * nl * nr;
    info_bytes_survey += 12.0 * nnz + 4.0 * (r + 1) + 8.0 * c * nr + 8.0 * r * nl + 8.0 * nl * nr;
    info_flops += 2.0 * nnz * nr + 2.0 * r * nl * nr;
  }
  cudaFree(d_cnt);
  cudaFree(d_flag);
}

int ProjectPlan::run(void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  for (const SpmmLaunch& L : spmm_launches) {
    if (!L.n_items) continue;
    switch (L.nt) {
      case 1: launch_spmm<1>(L, d_spmm_descs, s); break;
      case 2: launch_spmm<2>(L, d_spmm_descs, s); break;
      case 3: launch_spmm<3>(L, d_spmm_descs, s); break;
      case 4: launch_spmm<4>(L, d_spmm_descs, s); break;
      case 5: launch_spmm<5>(L, d_spmm_descs, s); break;
      case 6: launch_spmm<6>(L, d_spmm_descs, s); break;
      case 7: launch_spmm<7>(L, d_spmm_descs, s); break;
      default: launch_spmm<8>(L, d_spmm_descs, s); break;
    }
  }
  for (const ProjLaunch& L : launches) {
    if (!L.n_items) continue;
    if (L.has_a) dispatch_mt<true>(L, this, s); else dispatch_mt<false>(L, this, s);
  }
  LRBMS_CUDA_CHECK(ctx, cudaGetLastError());
  return LRBMS_OK;
}

}  // namespace

extern "C" {

int lrbms_spmm_plan_create(lrbms_handle_t h, int32_t n_desc, const lrbms_spmm_desc_t* descs_host, lrbms_plan_t* out) {
  LRBMS_REQUIRE(h, h && out && (n_desc == 0 || descs_host), "spmm_plan_create: null argument");
  LRBMS_CUDA_CHECK(h, cudaSetDevice(h->device));
  std::vector<lrbms_spmm_desc_t> descs(descs_host, descs_host + n_desc);
  for (const auto& d : descs) {
    LRBMS_REQUIRE(h, d.rowptr && d.colind && d.values && d.V && d.W, "spmm_plan_create: null pointer in descriptor");
    LRBMS_REQUIRE(h, d.n_rows >= 0 && d.n_cols >= 0 && d.N >= 1 && d.ldv >= d.N && d.ldw >= d.N,
                  "spmm_plan_create: inconsistent sizes in descriptor");
  }
  SpmmPlan* P = new SpmmPlan();
  P->ctx = h;
  P->kind = PLAN_SPMM;
  int rc = plan_upload(P, &P->d_descs, descs);
  std::vector<std::vector<SpmmItem>> by_nt;
  build_spmm_items(descs, h->sm_count, by_nt);
  for (int nt = 1; nt <= 8 && !rc; ++nt) {
    if (by_nt[nt - 1].empty()) continue;
    SpmmLaunch L;
    L.nt = nt;
    L.n_items = (int)by_nt[nt - 1].size();
    rc = plan_upload(P, &L.d_items, by_nt[nt - 1]);
    P->launches.push_back(L);
    P->info_launches += 1;
    P->info_ctas += L.n_items;
  }
  if (rc) { lrbms_plan_destroy(P); return rc; }
  // accounting: nnz from the device row pointers
  for (const auto& d : descs) {
    int32_t nnz = 0;
    if (d.n_rows > 0) cudaMemcpy(&nnz, d.rowptr + d.n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost);
    double b = 12.0 * nnz + 4.0 * (d.n_rows + 1) + 8.0 * (double)d.n_cols * d.N + 8.0 * (double)d.n_rows * d.N;
    P->info_bytes += b;
    P->info_bytes_survey += b;
    P->info_flops += 2.0 * nnz * d.N;
  }
  *out = P;
  return LRBMS_OK;
}

int lrbms_project_plan_create(lrbms_handle_t h, int32_t n_desc, const lrbms_project_desc_t* descs_host,
                              lrbms_plan_t* out) {
  LRBMS_REQUIRE(h, h && out && (n_desc == 0 || descs_host), "project_plan_create: null argument");
  LRBMS_CUDA_CHECK(h, cudaSetDevice(h->device));
  for (int i = 0; i < n_desc; ++i) {
    const auto& d = descs_host[i];
    LRBMS_REQUIRE(h, d.VL && d.VR && d.out, "project_plan_create: null pointer in descriptor");
    LRBMS_REQUIRE(h, d.rowptr == nullptr || (d.colind && d.values), "project_plan_create: rowptr without colind/values");
    LRBMS_REQUIRE(h, d.rowptr != nullptr || d.n_rows == d.n_cols, "project_plan_create: identity operator needs n_rows == n_cols");
    LRBMS_REQUIRE(h, d.NL >= 1 && d.NR >= 1 && d.ldl >= d.NL && d.ldr >= d.NR && d.ldo >= d.NR && d.n_rows >= 0,
                  "project_plan_create: inconsistent sizes in descriptor");
  }
  ProjectPlan* P = new ProjectPlan();
  P->ctx = h;
  P->kind = PLAN_PROJECT;
  int rc = 0;

  std::vector<DevDesc> dd(n_desc);
  std::vector<lrbms_spmm_desc_t> spmm_descs;
  // ---- pass 1: decide fused vs two-step, allocate SpMM scratch for the two-step descriptors
  for (int i = 0; i < n_desc && !rc; ++i) {
    const auto& d = descs_host[i];
    DevDesc& x = dd[i];
    x.rowptr = d.rowptr; x.colind = d.colind; x.values = d.values; x.n_rows = d.n_rows; x.n_cols = d.n_cols;
    x.VL = d.VL; x.ldl = d.ldl; x.NL = d.NL; x.VR = d.VR; x.ldr = d.ldr; x.NR = d.NR; x.out = d.out; x.ldo = d.ldo;
    x.alpha = d.alpha;
    const bool fits = d.NL <= 8 * kMaxTile && d.NR <= 8 * kMaxTile;
    if (d.rowptr && !fits) {
      // W = A VR into scratch, then G = VL^T W with the dense kernel
      double* W = nullptr;
      const int ldw = (d.NR + 3) & ~3;
      rc = plan_alloc(P, &W, (size_t)std::max(1, d.n_rows) * ldw);
      if (rc) break;
      lrbms_spmm_desc_t sd{d.rowptr, d.colind, d.values, d.n_rows, d.n_cols, d.VR, d.ldr, d.NR, W, ldw};
      spmm_descs.push_back(sd);
      x.rowptr = nullptr; x.colind = nullptr; x.values = nullptr; x.n_cols = d.n_rows;
      x.VR = W; x.ldr = ldw;
    }
  }
  if (rc) { lrbms_plan_destroy(P); return rc; }

  P->host_descs.assign(descs_host, descs_host + n_desc);

  // ---- pass 2: work items.  Output chunks of at most 40 x 40; rows split so the batch fills the machine.
  int64_t unit_rows = 0;
  for (int i = 0; i < n_desc; ++i) {
    const auto& x = dd[i];
    int64_t chunks = (int64_t)((x.NL + 8 * kMaxTile - 1) / (8 * kMaxTile)) * ((x.NR + 8 * kMaxTile - 1) / (8 * kMaxTile));
    unit_rows += chunks * x.n_rows;
  }
  const int64_t target_ctas = (int64_t)h->sm_count * 12;
  int64_t rows_per_cta = std::max<int64_t>(128, ((unit_rows / std::max<int64_t>(1, target_ctas) + 31) / 32) * 32);
  rows_per_cta = std::min<int64_t>(rows_per_cta, 2048);

  struct Key { int mt, nt; bool has_a; };
  std::vector<std::vector<ProjItem>> buckets(kMaxTile * kMaxTile * 2);
  std::vector<int64_t> group_base;
  int64_t n_partials = 0;   // partial slots, kPartialStride doubles each
  int32_t n_groups = 0;
  for (int i = 0; i < n_desc; ++i) {
    const auto& x = dd[i];
    const int n_lch = (x.NL + 8 * kMaxTile - 1) / (8 * kMaxTile), n_rch = (x.NR + 8 * kMaxTile - 1) / (8 * kMaxTile);
    // balanced chunk widths (e.g. 100 -> 3 chunks of 40, 32, 32 is worse than 3 x 5 tiles, 4, 4): use tiles
    const int lt = (x.NL + 7) / 8, rt = (x.NR + 7) / 8;
    const bool symmetric = descs_host[i].symmetric != 0 && x.NL == x.NR;
    for (int lc = 0; lc < n_lch; ++lc) {
      const int lt0 = (int)((int64_t)lt * lc / n_lch), lt1 = (int)((int64_t)lt * (lc + 1) / n_lch);
      for (int rc_ = 0; rc_ < n_rch; ++rc_) {
        if (symmetric && rc_ > lc) continue;          // upper chunks are mirrored from the lower ones
        const int rt0 = (int)((int64_t)rt * rc_ / n_rch), rt1 = (int)((int64_t)rt * (rc_ + 1) / n_rch);
        const int mt = lt1 - lt0, nt = rt1 - rt0;
        const int n_split = (int)std::max<int64_t>(1, (x.n_rows + rows_per_cta - 1) / rows_per_cta);
        const int64_t rows_each = ((((int64_t)x.n_rows + n_split - 1) / n_split) + 3) / 4 * 4;
        const int32_t group = n_groups++;
        group_base.push_back(n_partials);
        int slot = 0;
        std::vector<ProjItem>& bucket = buckets[((mt - 1) * kMaxTile + (nt - 1)) * 2 + (x.rowptr ? 1 : 0)];
        const size_t first = bucket.size();
        for (int64_t r0 = 0; r0 < std::max(1, x.n_rows); r0 += std::max<int64_t>(4, rows_each)) {
          ProjItem it;
          it.desc = i; it.l0 = 8 * lt0; it.c0 = 8 * rt0;
          it.row0 = (int32_t)r0; it.row1 = (int32_t)std::min<int64_t>(x.n_rows, r0 + std::max<int64_t>(4, rows_each));
          it.group = group; it.slot = slot++; it.group_size = 0;
          it.mirror = (symmetric && rc_ != lc) ? 1 : 0;
          bucket.push_back(it);
        }
        for (size_t k = first; k < bucket.size(); ++k) bucket[k].group_size = slot;
        if (slot > 1) n_partials += slot;
      }
    }
  }
  rc = plan_upload(P, &P->d_descs, dd);
  if (!rc) rc = plan_upload(P, &P->d_group_base, group_base);
  if (!rc) rc = plan_alloc(P, &P->d_partials, (size_t)std::max<int64_t>(1, n_partials) * kPartialStride);
  if (!rc) rc = plan_alloc(P, &P->d_flags, (size_t)std::max<int64_t>(1, n_partials));
  if (!rc) rc = plan_alloc(P, &P->d_counters, (size_t)std::max(1, n_groups));
  if (!rc) {
    cudaMemset(P->d_counters, 0, sizeof(int32_t) * std::max(1, n_groups));
    cudaMemset(P->d_flags, 0, sizeof(int32_t) * std::max<int64_t>(1, n_partials));
  }
  for (int mt = 1; mt <= kMaxTile && !rc; ++mt)
    for (int nt = 1; nt <= kMaxTile && !rc; ++nt)
      for (int ha = 0; ha < 2 && !rc; ++ha) {
        auto& bucket = buckets[((mt - 1) * kMaxTile + (nt - 1)) * 2 + ha];
        if (bucket.empty()) continue;
        ProjLaunch L;
        L.mt = mt; L.nt = nt; L.has_a = ha != 0; L.n_items = (int)bucket.size();
        rc = plan_upload(P, &L.d_items, bucket);
        P->launches.push_back(L);
        P->info_launches += 1;
        P->info_ctas += L.n_items;
      }
  if (!rc && !spmm_descs.empty()) {
    rc = plan_upload(P, &P->d_spmm_descs, spmm_descs);
    std::vector<std::vector<SpmmItem>> by_nt;
    build_spmm_items(spmm_descs, h->sm_count, by_nt);
    for (int nt = 1; nt <= 8 && !rc; ++nt) {
      if (by_nt[nt - 1].empty()) continue;
      SpmmLaunch L;
      L.nt = nt; L.n_items = (int)by_nt[nt - 1].size();
      rc = plan_upload(P, &L.d_items, by_nt[nt - 1]);
      P->spmm_launches.push_back(L);
      P->info_launches += 1;
    }
  }
  if (rc) { lrbms_plan_destroy(P); return rc; }
  *out = P;
  return LRBMS_OK;
}

}  // extern "C"
