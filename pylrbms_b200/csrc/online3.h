// solve_kernel_v3 (online3.cu): parameters and host entry points used by the online plan (online.cu).
#pragma once
#include "common.cuh"
#include "symbolic3.h"

struct V3Params {
  int32_t n_red, n_pad, ntc, np, Q, Qf, n_theta, n_a_tiles;
  int32_t n_win, acc_rows, n_partial, max_col;      // window tiles (the zero tile is n_win), accumulator rows, partial blocks, tiles per column
  int32_t region_doubles, back_stage_doubles;
  int32_t max_steps;         // early-update steps of one panel (peak): size of the staged step table
  const V3Own* own;          // [(np + 1) * kV3Warps]
  const V3Panel* pan;        // [np]
  const int4* steps;         // {a0, a1, b0, b1}
  const int4* ccol;          // [ntc] {first tile, tiles, 0, tile (J + 1, J) exists}
  const int32_t* row_idx;    // closed pattern
  const double* a_tiles;     // [Q][n_a_tiles][64] row-major operator tiles
  const double* rhs;         // [Qf][n_pad]
  int64_t work_stride;       // doubles of factor scratch per CTA: tiles of the closed pattern * 64
  long long* timing;         // developer builds (-DLRBMS_DEVTOOLS): [16 warps][8 phases] SM cycles of CTA 0, else NULL
};

size_t lrbms_v3_smem_bytes(const V3Params& P);
// raises the kernel's shared-memory limit; LRBMS_ERR_UNSUPPORTED (without an error message) when it does not fit
int lrbms_v3_prepare(lrbms_context* ctx, const V3Params& P, size_t* smem_out);
void lrbms_v3_launch(const V3Params& P, int grid, size_t smem, int64_t n_mu, const double* theta, double* u, int32_t* info,
                     double* work, cudaStream_t s);
int lrbms_v3_back_stages();
