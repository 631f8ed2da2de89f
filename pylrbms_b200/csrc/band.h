// Block-banded out-of-HBM Cholesky for large reduced systems (band.cu); used by the online plan (online.cu).
#pragma once
#include "common.cuh"

constexpr int kBandNB = 64;                       // block size (8 x 8 DMMA tiles)
constexpr int kBandMaxKb = 63;                    // block half bandwidth supported by the substitution kernel's ring
constexpr size_t kBandWorkspaceBudget = (size_t)32 << 30;   // default factor workspace per chunk of parameters (C4: all 64 in one)

struct lrbms_band_plan {
  int32_t n_red = 0, n_pad = 0, nbc = 0, kb = 0, Q = 0, Qf = 0, n_a = 0, half_bandwidth = 0;
  int64_t per_mu_doubles = 0, winv_off = 0;
  const int32_t* d_a_map = nullptr;
  const double* d_a_blocks = nullptr;
  const double* d_rhs = nullptr;
  double flops_per_mu = 0;
  size_t upd_smem = 0, trsm_smem = 0, potrf_smem = 0;
  // look-ahead: the diagonal block's factorisation runs on a side stream beside the off-diagonal updates of its column
  cudaStream_t side = nullptr;
  cudaEvent_t ev_diag = nullptr, ev_potrf = nullptr;
};
void lrbms_band_release(lrbms_band_plan& B);

// host_blocks: the packed reduced blocks (host copy); host_rhs: [Qf][n_red]
int lrbms_band_build(lrbms_plan* plan, lrbms_band_plan& B, int32_t n_sub, const int32_t* sizes, const int32_t* offsets, int32_t Q,
                     int32_t Qf, int32_t n_blocks, const int32_t* bi, const int32_t* bj, const int64_t* block_offset,
                     const double* host_blocks, const double* host_rhs);
size_t lrbms_band_workspace_bytes(const lrbms_band_plan& B, int64_t n_mu);
int64_t lrbms_band_chunk(const lrbms_band_plan& B, int64_t n_mu, size_t workspace_bytes);
int lrbms_band_solve(lrbms_context* ctx, const lrbms_band_plan& B, int64_t n_mu, const double* theta, double* u, int32_t* info,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);
