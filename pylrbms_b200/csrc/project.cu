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
  int32_t diag;       // gram_kernel: diagonal chunk of a symmetric descriptor (VL and VR chunks are the same columns)
  int32_t nl, nr;     // extent of the output chunk this item owns.  A gram CTA computes 2 x 2 quadrants of whole tiles and can
                      // cover more columns than its chunk has (7 tiles -> quadrants of 4 + 4); it must not store them: the
                      // neighbouring chunk owns those entries, and where that one is a *mirrored* chunk its values differ in
                      // the last bit (v_a^T A v_b vs v_b^T A v_a), which made the stored result depend on the write order
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

// ------------------------------------------------------------------------------------------------------
//  CSR walk in DMMA B-fragment layout
// ------------------------------------------------------------------------------------------------------
// one k-step (4 rows): b[n] = (A VR)[k0 + t][8n + g].  Lane g of row group t holds entry pb + g of its row (one
// coalesced load per 8 entries instead of 8 broadcast loads; the first chunk arrives preloaded in myc / myv) and the
// entries are handed round with shuffles, so the gathers of eight entries are independent and in flight together.
// Same summation order as a serial walk of the row.
template <int NT>
__device__ __forceinline__ void csr_row_fragment(const int32_t* __restrict__ ci, const double* __restrict__ va, int len,
                                                 int maxlen, int myc, double myv, const double* __restrict__ VR, int ldr,
                                                 int nr, int g, int t, double (&b)[NT]) {
#pragma unroll
  for (int n = 0; n < NT; ++n) b[n] = 0.0;
  for (int pb = 0; pb < maxlen; pb += 8) {
    if (pb > 0) {
      myc = 0;
      myv = 0.0;
      if (pb + g < len) { myc = ci[pb + g]; myv = va[pb + g]; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (pb + j >= maxlen) break;                       // warp-uniform
      const int src = 4 * j + t;
      const double av = __shfl_sync(0xffffffffu, myv, src);
      const int cj = __shfl_sync(0xffffffffu, myc, src);
      if (pb + j < len) {
        const double* __restrict__ vr = VR + (int64_t)cj * ldr;
#pragma unroll
        for (int n = 0; n < NT; ++n)
          if (8 * n + g < nr) b[n] = fma(av, vr[8 * n], b[n]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
//  combination of the CTAs that share one output chunk (fixed order -> bit-reproducible) and the store
// ------------------------------------------------------------------------------------------------------
// red: this CTA's sums in shared memory, kSlot doubles, element (a, bb) at idx(a, bb); slot_units: partial slots of
// kPartialStride doubles one CTA occupies
template <int kSlot, typename Idx>
__device__ __forceinline__ void combine_and_store(const ProjItem& it, const DevDesc& D, const double* red, bool cta_any,
                                                  int nl, int nr, double* __restrict__ partials, int32_t* __restrict__ flags,
                                                  int32_t* __restrict__ counters, const int64_t* __restrict__ group_partial_base,
                                                  int slot_units, int* s_last, Idx idx) {
  const int nthreads = blockDim.x;
  if (it.group_size == 1) {
    for (int e = threadIdx.x; e < nl * nr; e += nthreads) {
      const int a = e / nr, bb = e - a * nr;
      const double v = D.alpha * red[idx(a, bb)];
      D.out[(int64_t)(it.l0 + a) * D.ldo + it.c0 + bb] = v;
      if (it.mirror) D.out[(int64_t)(it.c0 + bb) * D.ldo + it.l0 + a] = v;
    }
    return;
  }
  const int64_t base = group_partial_base[it.group];
  double* my = partials + (base + (int64_t)it.slot * slot_units) * kPartialStride;
  if (cta_any)
    for (int e = threadIdx.x; e < kSlot; e += nthreads) my[e] = red[e];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    flags[base + it.slot] = cta_any ? 1 : 0;
    __threadfence();
    const int prev = atomicAdd(&counters[it.group], 1);
    *s_last = (prev == it.group_size - 1);
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  for (int e = threadIdx.x; e < nl * nr; e += nthreads) {
    const int a = e / nr, bb = e - a * nr;
    const int i = idx(a, bb);
    double s = 0.0;
    for (int sl = 0; sl < it.group_size; ++sl)
      if (__ldcg(&flags[base + sl])) s += __ldcg(&partials[(base + (int64_t)sl * slot_units) * kPartialStride + i]);
    D.out[(int64_t)(it.l0 + a) * D.ldo + it.c0 + bb] = D.alpha * s;
    if (it.mirror) D.out[(int64_t)(it.c0 + bb) * D.ldo + it.l0 + a] = D.alpha * s;
  }
  __syncthreads();
  if (threadIdx.x == 0) counters[it.group] = 0;   // self-cleaning for the next run
}

// ------------------------------------------------------------------------------------------------------
//  project_kernel: fused G = alpha VL^T (A VR) (HAS_A) and narrow dense G = alpha VL^T VR, one 8MT x 8NT chunk per CTA
// ------------------------------------------------------------------------------------------------------
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
    // Warp w owns the k-steps k0 = row0 + 4 (w + kWarps j).  32 of them are screened at once: lane j looks at the row
    // pointers of step j, a ballot gives the steps that have non-zeros at all (coupling blocks: interface rows only).
    // The row pointers, the first eight entries of each row and the VL fragment of the *next* live k-step are fetched
    // before the current one is processed, so only the gathers of VR rows remain on the dependent chain of a step.
    constexpr int kStep = 4 * kWarps;
    int kbase = it.row0 + 4 * warp - 32 * kStep;
    unsigned todo = 0;
    auto next_k = [&]() -> int {
      while (!todo) {
        kbase += 32 * kStep;
        if (kbase >= it.row1) return -1;
        const int kmine = kbase + lane * kStep;
        int has = 0;
        if (kmine < it.row1) has = D.rowptr[min(kmine + 4, it.row1)] != D.rowptr[kmine];
        todo = __ballot_sync(0xffffffffu, has);
      }
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      return kbase + j * kStep;
    };
    struct Step { int p0, len, myc; double myv; double a[MT]; };
    auto fetch = [&](int k0, Step& S) {
      const int row = k0 + t;
      const bool row_ok = row < it.row1;
      S.p0 = 0; S.len = 0; S.myc = 0; S.myv = 0.0;
      if (row_ok) {
        S.p0 = D.rowptr[row];
        S.len = D.rowptr[row + 1] - S.p0;
      }
      if (g < S.len) { S.myc = D.colind[S.p0 + g]; S.myv = D.values[S.p0 + g]; }
      const double* __restrict__ vl = VL + (int64_t)row * D.ldl;
#pragma unroll
      for (int m = 0; m < MT; ++m) S.a[m] = (row_ok && 8 * m + g < nl) ? vl[8 * m] : 0.0;
    };
    Step cur, nxt;
    int kcur = next_k();
    if (kcur >= 0) fetch(kcur, cur);
    while (kcur >= 0) {
      const int knext = next_k();
      if (knext >= 0) fetch(knext, nxt);
      const int maxlen = __reduce_max_sync(0xffffffffu, cur.len);
      double b[NT];
      csr_row_fragment<NT>(D.colind + cur.p0, D.values + cur.p0, cur.len, maxlen, cur.myc, cur.myv, VR, D.ldr, nr, g, t, b);
      any_work = true;
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma884(acc[m][n][0], acc[m][n][1], cur.a[m], b[n]);
      cur = nxt;
      kcur = knext;
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
  combine_and_store<MT * NT * 64>(it, D, red, s_any != 0, nl, nr, partials, flags, counters, group_partial_base, 1, &s_last,
                                  [](int a, int bb) { return ((a >> 3) * NT + (bb >> 3)) * 64 + (a & 7) * 8 + (bb & 7); });
}

// ------------------------------------------------------------------------------------------------------
//  gram_kernel: wide dense G = alpha VL^T VR (the estimator Grams: 100 ... 200 columns on both sides).
//
//  A 40 x 40 chunk per CTA needs 5 flop per byte fetched from L2 -- at the FP64 tensor rate that is ~7 TB/s of L2
//  bandwidth, which is what the one-warp-one-chunk kernel above runs into.  Here a CTA owns a (16 WM) x (16 WN) chunk
//  (up to 80 x 80): kGramRows rows of the VL and VR column blocks are staged once per CTA through a cp.async ring and
//  shared by four warps, one 8 WM x 8 WN quadrant each.  Two such CTAs are resident per SM, so one CTA's barriers,
//  pipeline fill and epilogue overlap the other's DMMAs.  Diagonal chunks of a pure Gram (VL and VR the same array) stage
//  their columns once and use them for both operands.
// ------------------------------------------------------------------------------------------------------
constexpr int kGramThreads = 128;    // four warps = four quadrants; two CTAs per SM hide each other's barriers and epilogues
constexpr int kGramRows = 16;        // rows per ring stage (4 k-steps)
constexpr int kGramStages = 4;
constexpr int kGramSlotUnits = 4;    // a gram partial slot = 4 project slots (100 tiles of 64 doubles)
template <int WM, int WN>
__host__ __device__ constexpr int gram_stride() { return 16 * WM + 16 * WN + 4; }   // = 4 mod 16: conflict-free fragment loads
template <int WM, int WN>
__host__ __device__ constexpr int gram_smem_doubles() {
  return (kGramStages * kGramRows * gram_stride<WM, WN>() > 4 * WM * WN * 64) ? kGramStages * kGramRows * gram_stride<WM, WN>()
                                                                               : 4 * WM * WN * 64;
}

template <int WM, int WN>
__global__ void __launch_bounds__(kGramThreads, 2)
gram_kernel(const ProjItem* __restrict__ items, const DevDesc* __restrict__ descs, double* __restrict__ partials,
            int32_t* __restrict__ flags, int32_t* __restrict__ counters, const int64_t* __restrict__ group_partial_base) {
  extern __shared__ __align__(16) double ring[];
  __shared__ int s_last;
  constexpr int S = gram_stride<WM, WN>();
  const ProjItem it = items[blockIdx.x];
  const DevDesc D = descs[it.desc];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int nl = min(it.nl, D.NL - it.l0), nr = min(it.nr, D.NR - it.c0);
  const bool diag = it.diag != 0;                       // VL and VR chunks are the same columns: stage them once
  const double* __restrict__ gL = D.VL + it.l0;
  const double* __restrict__ gR = D.VR + it.c0;
  // 16-byte copies need 16-byte aligned rows (column-slice views of a slab with an odd column offset are not)
  const bool al16 = ((((uintptr_t)gL) | ((uintptr_t)gR)) & 15) == 0 && (D.ldl & 1) == 0 && (D.ldr & 1) == 0;
  // columns that may be touched without leaving the row pitch (what lies between N and ld only feeds unwritten outputs)
  const int wl = min(16 * WM, D.ldl - it.l0), wr = min(16 * WN, D.ldr - it.c0);
  const int n_rows = it.row1 - it.row0;
  const int n_stage = (n_rows + kGramRows - 1) / kGramRows;

  // Staging: warp w copies rows 4w .. 4w + 3 of a stage; a lane takes the 16-byte (or 8-byte) chunks lane, lane + 32, ...
  // of each row.  The per-lane column offsets and byte counts do not depend on the stage and are set up once.
  constexpr int kChunks16 = 8 * WM + 8 * WN, kIter16 = (kChunks16 + 31) / 32;
  constexpr int kChunks8 = 16 * WM + 16 * WN, kIter8 = (kChunks8 + 31) / 32;
  const int n_left16 = 8 * WM, n_all16 = diag ? 8 * WM : kChunks16;
  const int n_left8 = 16 * WM, n_all8 = diag ? 16 * WM : kChunks8;
  auto issue = [&](int st) {
    if (st < n_stage) {
      double* dst = ring + (st % kGramStages) * (kGramRows * S) + 4 * warp * S;
      const int r0 = it.row0 + st * kGramRows + 4 * warp;
      if (al16) {
#pragma unroll
        for (int i = 0; i < kIter16; ++i) {
          const int cc = lane + 32 * i;
          if (cc < n_all16) {
            const bool left = cc < n_left16;
            const int col = 2 * (left ? cc : cc - n_left16);
            const int colbytes = max(0, min(2, (left ? wl : wr) - col)) * 8;
            const double* src0 = left ? gL + col : gR + col;
            const int64_t ld = left ? D.ldl : D.ldr;
            double* d0 = dst + (left ? col : 16 * WM + col);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int bytes = (r0 + r < it.row1) ? colbytes : 0;
              cp_async16_zfill(d0 + r * S, bytes ? src0 + (int64_t)(r0 + r) * ld : gL, bytes);
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < kIter8; ++i) {
          const int cc = lane + 32 * i;
          if (cc < n_all8) {
            const bool left = cc < n_left8;
            const int col = left ? cc : cc - n_left8;
            const bool col_ok = col < (left ? wl : wr);
            const double* src0 = left ? gL + col : gR + col;
            const int64_t ld = left ? D.ldl : D.ldr;
            double* d0 = dst + (left ? col : 16 * WM + col);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const bool ok = col_ok && r0 + r < it.row1;
              cp_async8_zfill(d0 + r * S, ok ? src0 + (int64_t)(r0 + r) * ld : gL, ok ? 8 : 0);
            }
          }
        }
      }
    }
    cp_async_commit();
  };

  double acc[WM][WN][2];
#pragma unroll
  for (int m = 0; m < WM; ++m)
#pragma unroll
    for (int n = 0; n < WN; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;

#pragma unroll
  for (int s = 0; s < kGramStages - 1; ++s) issue(s);
  const int aoff = wm * 8 * WM + g;
  const int boff = (diag ? 0 : 16 * WM) + wn * 8 * WN + g;
  for (int st = 0; st < n_stage; ++st) {
    cp_async_wait<kGramStages - 2>();
    __syncthreads();                       // stage st has landed for everybody; everybody is done with stage st - 1
    issue(st + kGramStages - 1);           // ... whose buffer is refilled now
    const double* base = ring + (st % kGramStages) * (kGramRows * S);
    double a[2][WM], b[2][WN];
    auto load = [&](int ks, double (&av)[WM], double (&bv)[WN]) {
      const double* rowp = base + (4 * ks + t) * S;
#pragma unroll
      for (int m = 0; m < WM; ++m) av[m] = rowp[aoff + 8 * m];
#pragma unroll
      for (int n = 0; n < WN; ++n) bv[n] = rowp[boff + 8 * n];
    };
    load(0, a[0], b[0]);
#pragma unroll
    for (int i = 0; i < kGramRows / 4; ++i) {            // the k-steps of this stage
      if (i + 1 < kGramRows / 4) load(i + 1, a[(i + 1) & 1], b[(i + 1) & 1]);
#pragma unroll
      for (int m = 0; m < WM; ++m)
#pragma unroll
        for (int n = 0; n < WN; ++n) dmma884(acc[m][n][0], acc[m][n][1], a[i & 1][m], b[i & 1][n]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();                         // the ring is dead: its memory becomes the reduction buffer

  // ---- red[quadrant][m][n][64]
  double* red = ring;
  const int q = wm * 2 + wn;
#pragma unroll
  for (int m = 0; m < WM; ++m)
#pragma unroll
    for (int n = 0; n < WN; ++n)
      *reinterpret_cast<double2*>(red + ((q * WM + m) * WN + n) * 64 + g * 8 + 2 * t) = make_double2(acc[m][n][0], acc[m][n][1]);
  __syncthreads();
  combine_and_store<4 * WM * WN * 64>(it, D, red, n_rows > 0, nl, nr, partials, flags, counters, group_partial_base,
                                      kGramSlotUnits, &s_last, [](int a, int bb) {
                                        const int qm = a / (8 * WM), qn = bb / (8 * WN);
                                        const int am = a - qm * 8 * WM, bn = bb - qn * 8 * WN;
                                        return (((qm * 2 + qn) * WM + (am >> 3)) * WN + (bn >> 3)) * 64 + (am & 7) * 8 + (bn & 7);
                                      });
}

// ------------------------------------------------------------------------------------------------------
//  SpMM  W = A V  in the same (row t, column 8n + g) lane layout; one warp per 4 rows and 8*NT columns.
//  (A one-lane-per-column variant that kept the V rows of a DG element's pattern in registers was tried and dropped:
//  three times the instructions per row at a quarter of the occupancy, 3.4 ms instead of 1.2 ms; 16-byte gathers (two
//  columns per lane) were slower too, 1.8 ms -- profiles/README.md.)
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
    const bool row_ok = row < it.row1;
    int p0 = 0, len = 0, myc = 0;
    double myv = 0.0;
    if (row_ok) {
      p0 = D.rowptr[row];
      len = D.rowptr[row + 1] - p0;
    }
    if (g < len) { myc = D.colind[p0 + g]; myv = D.values[p0 + g]; }
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    double b[NT];
    csr_row_fragment<NT>(D.colind + p0, D.values + p0, len, maxlen, myc, myv, V, D.ldv, nr, g, t, b);
    if (row_ok) {
      double* __restrict__ w = D.W + (int64_t)row * D.ldw + it.c0 + g;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (8 * n + g < nr) w[8 * n] = b[n];
    }
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

static void dispatch_spmm(const SpmmLaunch& L, const lrbms_spmm_desc_t* d_descs, cudaStream_t s) {
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

// launches are kept sorted by size (largest first) and dealt round-robin to the caller's stream and the side streams
int SpmmPlan::run(void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int n_live = 0;
  for (const SpmmLaunch& L : launches) n_live += L.n_items > 0;
  const int ns = n_live > 1 ? ctx_streams(ctx) : 1;
  if (ns > 1) ctx_fork(ctx, s);
  int k = 0;
  for (const SpmmLaunch& L : launches) {
    if (!L.n_items) continue;
    dispatch_spmm(L, d_descs, ctx_stream(ctx, s, k++, ns));
  }
  if (ns > 1) ctx_join(ctx, s);
  LRBMS_CUDA_CHECK(ctx, cudaGetLastError());
  return LRBMS_OK;
}

// build the SpMM work list for a set of descriptors; shared by the SpMM plan and by the two-step projection
static void build_spmm_items(const std::vector<lrbms_spmm_desc_t>& descs, int sm_count,
                             std::vector<std::vector<SpmmItem>>& by_nt /* index nt-1 */) {
  by_nt.assign(8, {});
  // rows per CTA: aim at >= 8 CTAs per SM over the whole batch, in multiples of 32 rows (one k-step per warp)
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
  int mt, nt;         // tiles per CTA chunk (project_kernel) or per warp quadrant (gram_kernel)
  bool has_a;
  bool gram = false;
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

// raises the dynamic shared-memory limit of one gram_kernel instantiation; called once per launch bucket at plan creation
template <int WM, int WN>
static cudaError_t prepare_gram() {
  return cudaFuncSetAttribute(gram_kernel<WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(sizeof(double) * gram_smem_doubles<WM, WN>()));
}
template <int WM>
static cudaError_t prepare_gram_n(int nt) {
  switch (nt) {
    case 1: return prepare_gram<WM, 1>();
    case 2: return prepare_gram<WM, 2>();
    case 3: return prepare_gram<WM, 3>();
    case 4: return prepare_gram<WM, 4>();
    default: return prepare_gram<WM, 5>();
  }
}
static cudaError_t prepare_gram_mn(int mt, int nt) {
  switch (mt) {
    case 1: return prepare_gram_n<1>(nt);
    case 2: return prepare_gram_n<2>(nt);
    case 3: return prepare_gram_n<3>(nt);
    case 4: return prepare_gram_n<4>(nt);
    default: return prepare_gram_n<5>(nt);
  }
}

template <int WM, int WN>
static void launch_gram(const ProjLaunch& L, const ProjectPlan* P, cudaStream_t s) {
  const size_t smem = sizeof(double) * gram_smem_doubles<WM, WN>();
  gram_kernel<WM, WN><<<L.n_items, kGramThreads, smem, s>>>(L.d_items, P->d_descs, P->d_partials, P->d_flags, P->d_counters,
                                                        P->d_group_base);
}

template <int WM>
static void dispatch_gram_n(const ProjLaunch& L, const ProjectPlan* P, cudaStream_t s) {
  switch (L.nt) {
    case 1: launch_gram<WM, 1>(L, P, s); break;
    case 2: launch_gram<WM, 2>(L, P, s); break;
    case 3: launch_gram<WM, 3>(L, P, s); break;
    case 4: launch_gram<WM, 4>(L, P, s); break;
    default: launch_gram<WM, 5>(L, P, s); break;
  }
}

static void dispatch_gram(const ProjLaunch& L, const ProjectPlan* P, cudaStream_t s) {
  switch (L.mt) {
    case 1: dispatch_gram_n<1>(L, P, s); break;
    case 2: dispatch_gram_n<2>(L, P, s); break;
    case 3: dispatch_gram_n<3>(L, P, s); break;
    case 4: dispatch_gram_n<4>(L, P, s); break;
    default: dispatch_gram_n<5>(L, P, s); break;
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
  // any CUDA failure leaves the counts at NaN (never a silently wrong figure) and records the message
  auto give_up = [&](cudaError_t e) {
    info_bytes = info_bytes_survey = info_flops = NAN;
    lrbms_fail(ctx, LRBMS_ERR_CUDA, std::string("plan accounting: ") + cudaGetErrorString(e));
    if (d_cnt) cudaFree(d_cnt);
    if (d_flag) cudaFree(d_flag);
  };
  cudaError_t e = cudaMalloc((void**)&d_cnt, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_flag, (size_t)max_cols);
  if (e != cudaSuccess) { give_up(e); return; }
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
      e = cudaMemcpy(&nnz, d.rowptr + d.n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemset(d_cnt, 0, 2 * sizeof(unsigned long long));
      if (e == cudaSuccess) e = cudaMemset(d_flag, 0, (size_t)std::max(1, d.n_cols));
      if (e == cudaSuccess) {
        csr_stats_kernel<<<64, 256>>>(d.rowptr, d.colind, d.n_rows, d_flag, d_cnt);
        count_flags_kernel<<<64, 256>>>(d_flag, d.n_cols, d_cnt);
        e = cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost);
      }
      if (e != cudaSuccess) { give_up(e); return; }
    }
    info_bytes += 12.0 * nnz + 4.0 * (r + 1) + 8.0 * (double)cnt[1] * nr + 8.0 * (double)cnt[0] * nl + 8.0 * nl * nr;
    info_bytes_survey += 12.0 * nnz + 4.0 * (r + 1) + 8.0 * c * nr + 8.0 * r * nl + 8.0 * nl * nr;
    info_flops += 2.0 * nnz * nr + 2.0 * r * nl * nr;
  }
  cudaFree(d_cnt);
  cudaFree(d_flag);
}

// Phase 1: the scratch SpMMs of the two-step descriptors and the fused projections (independent of each other);
// phase 2: the dense projections, which read the SpMM scratch.  Inside a phase the launches are spread over streams.
int ProjectPlan::run(void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int n_live = 0;
  for (const SpmmLaunch& L : spmm_launches) n_live += L.n_items > 0;
  for (const ProjLaunch& L : launches) n_live += L.n_items > 0;
  const int ns = n_live > 1 ? ctx_streams(ctx) : 1;
  if (ns > 1) ctx_fork(ctx, s);
  int k = 0;
  bool any_spmm = false;
  for (const SpmmLaunch& L : spmm_launches) {
    if (!L.n_items) continue;
    any_spmm = true;
    dispatch_spmm(L, d_spmm_descs, ctx_stream(ctx, s, k++, ns));
  }
  for (const ProjLaunch& L : launches)
    if (L.n_items && L.has_a) dispatch_mt<true>(L, this, ctx_stream(ctx, s, k++, ns));
  if (ns > 1 && any_spmm) {
    ctx_join(ctx, s);
    ctx_fork(ctx, s);
    k = 0;
  }
  for (const ProjLaunch& L : launches) {
    if (!L.n_items || L.has_a) continue;
    if (L.gram) dispatch_gram(L, this, ctx_stream(ctx, s, k++, ns));
    else dispatch_mt<false>(L, this, ctx_stream(ctx, s, k++, ns));
  }
  if (ns > 1) ctx_join(ctx, s);
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
  std::stable_sort(P->launches.begin(), P->launches.end(),
                   [](const SpmmLaunch& a, const SpmmLaunch& b) { return (int64_t)a.n_items * a.nt > (int64_t)b.n_items * b.nt; });
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
  if (cudaStreamSynchronize((cudaStream_t)0) != cudaSuccess) {
    lrbms_plan_destroy(P);
    return lrbms_fail(h, LRBMS_ERR_CUDA, "spmm_plan_create: stream sync failed");
  }
  *out = P;
  return LRBMS_OK;
}

static inline bool two_step(const lrbms_project_desc_t& d) {
  return d.rowptr && !(d.NL <= 8 * kMaxTile && d.NR <= 8 * kMaxTile);
}
static inline size_t scratch_doubles(const lrbms_project_desc_t& d) {
  const size_t n = (size_t)std::max(1, d.n_rows) * (size_t)((d.NR + 3) & ~3);
  return (n + 31) & ~(size_t)31;
}

int lrbms_project_plan_scratch_bytes(int32_t n_desc, const lrbms_project_desc_t* descs_host, size_t* bytes) {
  if (!bytes || n_desc < 0 || (n_desc && !descs_host)) return LRBMS_ERR_INVALID;
  size_t n = 0;
  for (int i = 0; i < n_desc; ++i)
    if (two_step(descs_host[i])) n += scratch_doubles(descs_host[i]);
  *bytes = n * sizeof(double);
  return LRBMS_OK;
}

int lrbms_project_plan_create(lrbms_handle_t h, int32_t n_desc, const lrbms_project_desc_t* descs_host,
                              lrbms_plan_t* out) {
  return lrbms_project_plan_create_ws(h, n_desc, descs_host, nullptr, 0, 0, out);
}

int lrbms_project_plan_create_ws(lrbms_handle_t h, int32_t n_desc, const lrbms_project_desc_t* descs_host, void* scratch,
                                 size_t scratch_bytes, int64_t unit_rows_hint, lrbms_plan_t* out) {
  LRBMS_REQUIRE(h, h && out && (n_desc == 0 || descs_host), "project_plan_create: null argument");
  if (scratch) {
    size_t need = 0;
    lrbms_project_plan_scratch_bytes(n_desc, descs_host, &need);
    LRBMS_REQUIRE(h, scratch_bytes >= need && ((uintptr_t)scratch & 255) == 0,
                  "project_plan_create_ws: scratch too small or misaligned (see lrbms_project_plan_scratch_bytes)");
  }
  size_t scratch_pos = 0;
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
    if (two_step(d)) {
      // W = A VR into scratch, then G = VL^T W with the dense kernel
      double* W = nullptr;
      const int ldw = (d.NR + 3) & ~3;
      if (scratch) {
        W = reinterpret_cast<double*>(scratch) + scratch_pos;
        scratch_pos += scratch_doubles(d);
      } else {
        rc = plan_alloc(P, &W, (size_t)std::max(1, d.n_rows) * ldw);
        if (rc) break;
      }
      lrbms_spmm_desc_t sd{d.rowptr, d.colind, d.values, d.n_rows, d.n_cols, d.VR, d.ldr, d.NR, W, ldw};
      spmm_descs.push_back(sd);
      x.rowptr = nullptr; x.colind = nullptr; x.values = nullptr; x.n_cols = d.n_rows;
      x.VR = W; x.ldr = ldw;
    }
  }
  if (rc) { lrbms_plan_destroy(P); return rc; }

  P->host_descs.assign(descs_host, descs_host + n_desc);

  // ---- pass 2: work items.  Output chunks of at most 40 x 40 (project_kernel) or 80 x 80 (gram_kernel: dense with more
  //      than 40 columns on both sides); rows split so the batch fills the machine.
  auto is_gram = [&](const DevDesc& x) { return !x.rowptr && x.NL > 8 * kMaxTile && x.NR > 8 * kMaxTile; };
  int64_t unit_rows = 0;
  for (int i = 0; i < n_desc; ++i) {
    const auto& x = dd[i];
    const int cw = is_gram(x) ? 16 * kMaxTile : 8 * kMaxTile;
    int64_t chunks = (int64_t)((x.NL + cw - 1) / cw) * ((x.NR + cw - 1) / cw);
    unit_rows += chunks * x.n_rows * (is_gram(x) ? 3 : 1);     // a gram CTA does four times the work with one CTA per SM
  }
  if (unit_rows_hint > 0) unit_rows = unit_rows_hint;
  const int64_t target_ctas = (int64_t)h->sm_count * 12;
  int64_t rows_per_cta = std::max<int64_t>(128, ((unit_rows / std::max<int64_t>(1, target_ctas) + 31) / 32) * 32);
  rows_per_cta = std::min<int64_t>(rows_per_cta, 2048);

  std::vector<std::vector<ProjItem>> buckets(kMaxTile * kMaxTile * 3);   // [(mt, nt)][dense, fused, gram]
  std::vector<int64_t> group_base;
  int64_t n_partials = 0;   // partial slots, kPartialStride doubles each
  int32_t n_groups = 0;
  for (int i = 0; i < n_desc; ++i) {
    const auto& x = dd[i];
    const bool gram = is_gram(x);
    const int max_ct = gram ? 2 * kMaxTile : kMaxTile;           // tiles per chunk
    // balanced chunk widths (e.g. 100 -> 3 chunks of 40, 32, 32 is worse than 3 x 5 tiles, 4, 4): use tiles
    const int lt = (x.NL + 7) / 8, rt = (x.NR + 7) / 8;
    const int n_lch = (lt + max_ct - 1) / max_ct, n_rch = (rt + max_ct - 1) / max_ct;
    const bool symmetric = descs_host[i].symmetric != 0 && x.NL == x.NR;
    const int row_quant = gram ? kGramRows : 4;
    for (int lc = 0; lc < n_lch; ++lc) {
      const int lt0 = (int)((int64_t)lt * lc / n_lch), lt1 = (int)((int64_t)lt * (lc + 1) / n_lch);
      for (int rc_ = 0; rc_ < n_rch; ++rc_) {
        if (symmetric && rc_ > lc) continue;          // upper chunks are mirrored from the lower ones
        const int rt0 = (int)((int64_t)rt * rc_ / n_rch), rt1 = (int)((int64_t)rt * (rc_ + 1) / n_rch);
        // project_kernel: tiles per CTA; gram_kernel: tiles per warp quadrant (two quadrants per direction)
        const int mt = gram ? (lt1 - lt0 + 1) / 2 : lt1 - lt0, nt = gram ? (rt1 - rt0 + 1) / 2 : rt1 - rt0;
        const int64_t rpc = rows_per_cta;
        const int n_split = (int)std::max<int64_t>(1, (x.n_rows + rpc - 1) / rpc);
        const int64_t rows_each = std::max<int64_t>(
            row_quant, ((((int64_t)x.n_rows + n_split - 1) / n_split) + row_quant - 1) / row_quant * row_quant);
        const int32_t group = n_groups++;
        group_base.push_back(n_partials);
        int slot = 0;
        std::vector<ProjItem>& bucket = buckets[((mt - 1) * kMaxTile + (nt - 1)) * 3 + (gram ? 2 : (x.rowptr ? 1 : 0))];
        const size_t first = bucket.size();
        for (int64_t r0 = 0; r0 < std::max(1, x.n_rows); r0 += rows_each) {
          ProjItem it;
          it.desc = i; it.l0 = 8 * lt0; it.c0 = 8 * rt0;
          it.row0 = (int32_t)r0; it.row1 = (int32_t)std::min<int64_t>(x.n_rows, r0 + rows_each);
          it.group = group; it.slot = slot++; it.group_size = 0;
          it.mirror = (symmetric && rc_ != lc) ? 1 : 0;
          it.diag = (gram && symmetric && rc_ == lc && x.VL == x.VR && x.ldl == x.ldr) ? 1 : 0;
          it.nl = 8 * (lt1 - lt0); it.nr = 8 * (rt1 - rt0);
          bucket.push_back(it);
        }
        for (size_t k = first; k < bucket.size(); ++k) bucket[k].group_size = slot;
        if (slot > 1) n_partials += (int64_t)slot * (gram ? kGramSlotUnits : 1);
      }
    }
  }
  rc = plan_upload(P, &P->d_descs, dd);
  if (!rc) rc = plan_upload(P, &P->d_group_base, group_base);
  if (!rc) rc = plan_alloc(P, &P->d_partials, (size_t)std::max<int64_t>(1, n_partials) * kPartialStride);
  if (!rc) rc = plan_alloc(P, &P->d_flags, (size_t)std::max<int64_t>(1, n_partials));
  if (!rc) rc = plan_alloc(P, &P->d_counters, (size_t)std::max(1, n_groups));
  if (!rc) {
    cudaError_t e = cudaMemset(P->d_counters, 0, sizeof(int32_t) * std::max(1, n_groups));
    if (e == cudaSuccess) e = cudaMemset(P->d_flags, 0, sizeof(int32_t) * std::max<int64_t>(1, n_partials));
    if (e != cudaSuccess) rc = lrbms_fail(h, LRBMS_ERR_CUDA, std::string("project_plan_create: ") + cudaGetErrorString(e));
  }
  for (int mt = 1; mt <= kMaxTile && !rc; ++mt)
    for (int nt = 1; nt <= kMaxTile && !rc; ++nt)
      for (int kind = 0; kind < 3 && !rc; ++kind) {
        auto& bucket = buckets[((mt - 1) * kMaxTile + (nt - 1)) * 3 + kind];
        if (bucket.empty()) continue;
        ProjLaunch L;
        L.mt = mt; L.nt = nt; L.has_a = kind == 1; L.gram = kind == 2; L.n_items = (int)bucket.size();
        rc = plan_upload(P, &L.d_items, bucket);
        if (!rc && L.gram) {
          const cudaError_t e = prepare_gram_mn(mt, nt);
          if (e != cudaSuccess) rc = lrbms_fail(h, LRBMS_ERR_CUDA, std::string("gram_kernel shared-memory limit: ") + cudaGetErrorString(e));
        }
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
  std::stable_sort(P->spmm_launches.begin(), P->spmm_launches.end(),
                   [](const SpmmLaunch& a, const SpmmLaunch& b) { return (int64_t)a.n_items * a.nt > (int64_t)b.n_items * b.nt; });
  std::stable_sort(P->launches.begin(), P->launches.end(), [](const ProjLaunch& a, const ProjLaunch& b) {
    return (int64_t)a.n_items * a.mt * a.nt * (a.gram ? 4 : 1) > (int64_t)b.n_items * b.mt * b.nt * (b.gram ? 4 : 1);
  });
  // uploads and clears ran on the legacy default stream: finish them before a caller's non-blocking stream runs the plan
  if (!rc && cudaStreamSynchronize((cudaStream_t)0) != cudaSuccess) rc = lrbms_fail(h, LRBMS_ERR_CUDA, "project_plan_create: stream sync failed");
  if (rc) { lrbms_plan_destroy(P); return rc; }
  *out = P;
  return LRBMS_OK;
}

}  // extern "C"
