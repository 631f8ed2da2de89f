// Host-only schedule of solve_kernel_v3: the 8x8-tile sparse Cholesky of the reduced system processed in PANELS of two tile
// columns, so that every update step multiplies a 2 x 2 block of target tiles by two operand tiles from each side
// (8 DMMAs per 4 shared-memory fragment loads; the one-column kernel needs 2 loads per 2 DMMAs and is bound by the
// shared-memory port).  See online3.cu for the kernel; tests/test_symbolic3_emulator.py executes exactly these tables
// in NumPy against a dense Cholesky.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "symbolic.h"

constexpr int kV3Warps = 16;        // warp 0: critical chain; warps 1..15: update warps
constexpr int kV3MaxChunks = 4;     // early-update chunks per (panel, warp)
constexpr int kV3MaxFold = 4;       // partial blocks folded into one target block

// per target panel q and update warp w (1..15): what the warp owns.  Rows are tile rows; -1 = none.
struct V3Own {
  int32_t row[2];          // tile rows Ia, Ib of the warp's block in panel q (row[0] == -2: the right-hand-side block)
  int32_t acc[2];          // accumulator-buffer row index of each row (acc tiles acc[r] * 2 + column)
  int32_t prev[2];         // 1: the row also exists in panel q - 1 (its two tiles of that panel are solved by this warp)
  int32_t wprev[2][2];     // window slots of tiles (row, c0p), (row, c1p) of panel q - 1 (written by this warp), -1: none
  int32_t gprev[2][2];     // slots of the same tiles in the stored factor (global), -1: none
  int32_t amap[2][2];      // operator tiles of the targets (row, t0), (row, t1), -1: none
  int32_t exists[2][2];    // 1: target tile (row, t) is part of the (closed) pattern
  int32_t fold[kV3MaxFold];// partial blocks (index into the partial buffer) added to this block before the late update, -1: none
  int32_t n_chunks;
  int32_t chunk_step[kV3MaxChunks];   // first step of each early-update chunk
  int32_t chunk_n[kV3MaxChunks];      // steps in the chunk
  int32_t chunk_dest[kV3MaxChunks];   // -1: this warp's own block (kept in registers); >= 0: partial-buffer block
  int32_t chunk_kind[kV3MaxChunks];   // 0: tile block (4 operand slots per step); 1: right-hand-side block
  int32_t pad;                        // record size: 44 words = 176 bytes (16-byte multiples for the staging copies)
};

// per panel p: the diagonal block and the head rows (what the chain warp handles)
struct V3Panel {
  int32_t c0, c1;              // tile columns
  int32_t g_d00, g_d10, g_d11; // factor slots: W00 = L00^-1, L10, W11 = L11^-1
  int32_t w_d10;               // window slot of L10
  int32_t a_d00, a_d10, a_d11; // operator tiles of the diagonal block
  int32_t fold[kV3MaxFold];    // partial blocks of the diagonal block's early updates
  // head rows: rows c0, c1 of THIS panel as off-diagonal rows of panel p - 1 (solved by the chain warp)
  int32_t head_prev;           // 1: panel p - 1 exists and holds rows c0 / c1
  int32_t head_exists[2];      // row c0 / c1 present in panel p - 1
  int32_t head_acc[2];         // accumulator rows
  int32_t head_w[2][2];        // window slots of (c0, c0p), (c0, c1p), (c1, c0p), (c1, c1p)
  int32_t head_g[2][2];        // factor slots of the same
  int32_t acc_rows[2];         // accumulator rows of c0, c1 as diagonal rows (where the diagonal block lives)
  int32_t step0, n_steps;      // the early-update steps of this panel's blocks: steps[step0 .. step0 + n_steps)
  int32_t pad[2];              // record size: 32 words = 128 bytes (16-byte multiples for the staging copies)
};

struct lrbms_symbolic3 {
  bool ok = false;
  std::string why;             // why the panel schedule does not apply (ok == false)
  int32_t ntc = 0, np = 0, n_pad = 0;
  // closed tile pattern of L (both columns of a panel have the same off-diagonal rows), diagonal tile first
  std::vector<int32_t> col_ptr, row_idx, a_map, win_slot;
  int32_t n_win_slots = 0;     // live window tiles (peak); slot n_win_slots is an all-zero tile
  int32_t acc_rows = 0;        // rows of the accumulator ring
  int32_t n_partial = 0;       // partial blocks (peak per panel)
  int32_t max_steps = 0;       // early-update steps of one panel (peak)
  std::vector<V3Own> own;      // [(np + 1) * kV3Warps]
  std::vector<V3Panel> pan;    // [np]
  std::vector<int32_t> steps;  // 4 int32 per step: a0, a1, b0, b1 (window slots; right-hand-side block: K, unused, b0, b1)
  int64_t flops = 0;           // DMMA flops per parameter incl. the padding of 2 x 2 blocks
  int64_t n_tiles() const { return (int64_t)row_idx.size(); }
};

// builds the schedule from the tile pattern of `S` (which must have been built with lrbms_symbolic_build)
int lrbms_symbolic3_build(lrbms_symbolic3& S3, const lrbms_symbolic& S);
