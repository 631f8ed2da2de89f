// Host-only schedule of solve_kernel_v3 (see symbolic3.h).  No CUDA calls in this file.
#include "symbolic3.h"

#include <algorithm>
#include <map>
#include <numeric>

#include "../../include/lrbms_sm100.h"

namespace {
struct Chunk { int block, step0, n, kind; };   // block: -1 diagonal block, else index into the panel's owner blocks
}

int lrbms_symbolic3_build(lrbms_symbolic3& S3, const lrbms_symbolic& S) {
  auto no = [&](const char* why) { S3.ok = false; S3.why = why; return (int)LRBMS_OK; };
  S3 = lrbms_symbolic3();
  if (S.col_ptr.empty()) return no("tile schedule not built");
  if (S.ntc % 2) return no("odd number of tile columns");
  const int ntc = S.ntc, np = ntc / 2;
  S3.ntc = ntc; S3.np = np; S3.n_pad = S.n_pad;

  // ---- closed pattern: both columns of a panel get the union of their off-diagonal rows (explicit zero tiles)
  std::vector<std::vector<int32_t>> rows(np);                 // R_q: off-diagonal rows of panel q (>= 2q + 2), ascending
  for (int q = 0; q < np; ++q) {
    std::vector<int32_t>& r = rows[q];
    for (int c = 2 * q; c <= 2 * q + 1; ++c)
      for (int32_t s = S.col_ptr[c]; s < S.col_ptr[c + 1]; ++s)
        if (S.row_idx[s] >= 2 * q + 2) r.push_back(S.row_idx[s]);
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
  }
  // every row of panel q - 1 other than the diagonal rows of panel q must also be a row of panel q (its tiles of panel
  // q - 1 are solved by the warp that owns the row in panel q)
  for (int q = 1; q < np; ++q)
    for (int32_t I : rows[q - 1])
      if (I >= 2 * q + 2 && !std::binary_search(rows[q].begin(), rows[q].end(), I)) return no("a tile row leaves the pattern and comes back");
  S3.col_ptr.assign(ntc + 1, 0);
  for (int c = 0; c < ntc; ++c) {
    const int q = c / 2;
    S3.col_ptr[c + 1] = S3.col_ptr[c] + 1 + ((c % 2 == 0) ? 1 : 0) + (int32_t)rows[q].size();   // diag [+ (c1, c0)] + rows
    S3.row_idx.push_back(c);
    if (c % 2 == 0) S3.row_idx.push_back(c + 1);
    S3.row_idx.insert(S3.row_idx.end(), rows[q].begin(), rows[q].end());
  }
  const int64_t n_tiles = S3.n_tiles();
  auto slot3 = [&](int I, int c) -> int32_t {                  // factor slot of tile (I, c) in the closed pattern, -1: none
    const int32_t* b = S3.row_idx.data() + S3.col_ptr[c];
    const int32_t* e = S3.row_idx.data() + S3.col_ptr[c + 1];
    const int32_t* it = std::lower_bound(b, e, I);
    return (it != e && *it == I) ? (int32_t)(it - S3.row_idx.data()) : -1;
  };
  auto slot_orig = [&](int I, int c) -> int32_t {
    const int32_t* b = S.row_idx.data() + S.col_ptr[c];
    const int32_t* e = S.row_idx.data() + S.col_ptr[c + 1];
    const int32_t* it = std::lower_bound(b, e, I);
    return (it != e && *it == I) ? (int32_t)(it - S.row_idx.data()) : -1;
  };
  S3.a_map.assign(n_tiles, -1);
  for (int c = 0; c < ntc; ++c)
    for (int32_t s = S3.col_ptr[c]; s < S3.col_ptr[c + 1]; ++s) {
      const int32_t so = slot_orig(S3.row_idx[s], c);
      if (so >= 0) S3.a_map[s] = S.a_map[so];
    }

  // ---- window slots of the off-diagonal tiles (I, c), I >= 2 (c / 2) + 2: written in iteration c / 2, last read in
  //      iteration I / 2 - 1; the tiles written in iteration p may reuse the slots of all tiles whose row panel is <= p
  S3.win_slot.assign(n_tiles, -1);
  {
    std::vector<std::vector<int32_t>> by_row_panel(np);
    std::vector<int32_t> free_slots;
    int32_t n_slots = 0;
    for (int p = 0; p < np; ++p) {
      for (int32_t s : by_row_panel[p]) free_slots.push_back(S3.win_slot[s]);
      for (int c = 2 * p; c <= 2 * p + 1; ++c)
        for (int32_t I : rows[p]) {
          const int32_t s = slot3(I, c);
          if (!free_slots.empty()) { S3.win_slot[s] = free_slots.back(); free_slots.pop_back(); }
          else S3.win_slot[s] = n_slots++;
          by_row_panel[I / 2].push_back(s);
        }
    }
    S3.n_win_slots = n_slots;
  }
  const int32_t Z = S3.n_win_slots;                            // the all-zero tile
  auto wslot = [&](int I, int c) -> int32_t { const int32_t s = slot3(I, c); return s >= 0 ? S3.win_slot[s] : -1; };

  // ---- accumulator ring
  int span = 4;
  for (int q = 0; q < np; ++q)
    if (!rows[q].empty()) span = std::max(span, rows[q].back() - 2 * q + 1);
  S3.acc_rows = 4;
  while (S3.acc_rows < span + 2) S3.acc_rows *= 2;
  auto acc_of = [&](int I) { return I % S3.acc_rows; };

  // ---- panels
  S3.pan.assign(np, V3Panel());
  for (int p = 0; p < np; ++p) {
    V3Panel& P = S3.pan[p];
    P.c0 = 2 * p; P.c1 = 2 * p + 1;
    P.g_d00 = slot3(P.c0, P.c0); P.g_d10 = slot3(P.c1, P.c0); P.g_d11 = slot3(P.c1, P.c1);
    P.w_d10 = -1;
    P.a_d00 = S3.a_map[P.g_d00]; P.a_d10 = S3.a_map[P.g_d10]; P.a_d11 = S3.a_map[P.g_d11];
    for (int k = 0; k < kV3MaxFold; ++k) P.fold[k] = -1;
    P.head_prev = p > 0;
    P.acc_rows[0] = acc_of(P.c0); P.acc_rows[1] = acc_of(P.c1);
    P.step0 = P.n_steps = 0; P.pad[0] = P.pad[1] = 0;
    for (int r = 0; r < 2; ++r) {
      const int I = 2 * p + r;
      P.head_exists[r] = p > 0 && std::binary_search(rows[p - 1].begin(), rows[p - 1].end(), I);
      P.head_acc[r] = acc_of(I);
      for (int c = 0; c < 2; ++c) {
        P.head_w[r][c] = P.head_exists[r] ? wslot(I, 2 * (p - 1) + c) : -1;
        P.head_g[r][c] = P.head_exists[r] ? slot3(I, 2 * (p - 1) + c) : -1;
      }
    }
  }

  // ---- per target panel q = 0 .. np (q = np: only the right-hand-side block, for the solves of panel np - 1)
  S3.own.assign((size_t)(np + 1) * kV3Warps, V3Own());
  for (auto& o : S3.own) {
    o.row[0] = o.row[1] = -1;
    o.n_chunks = 0;
    for (int k = 0; k < kV3MaxFold; ++k) o.fold[k] = -1;
    for (int r = 0; r < 2; ++r) {
      o.acc[r] = 0; o.prev[r] = 0;
      for (int c = 0; c < 2; ++c) { o.wprev[r][c] = o.gprev[r][c] = o.amap[r][c] = -1; o.exists[r][c] = 0; }
    }
  }
  S3.n_partial = 0;
  S3.flops = 0;
  for (int q = 0; q <= np; ++q) {
    const int p = q - 1;
    const int t0 = 2 * q, t1 = 2 * q + 1;
    const std::vector<int32_t> none;
    const std::vector<int32_t>& R = q < np ? rows[q] : none;
    const int n_blocks = (int)(R.size() + 1) / 2;              // tile blocks; the right-hand-side block comes last
    if (n_blocks + 1 > kV3Warps - 1) return no("more target blocks per panel than update warps");
    // --- owners
    for (int b = 0; b <= n_blocks; ++b) {
      V3Own& o = S3.own[(size_t)q * kV3Warps + 1 + b];
      if (b == n_blocks) { o.row[0] = -2; o.prev[0] = p >= 0; continue; }
      for (int r = 0; r < 2; ++r) {
        if (2 * b + r >= (int)R.size()) continue;
        const int I = R[2 * b + r];
        o.row[r] = I;
        o.acc[r] = acc_of(I);
        o.prev[r] = p >= 0 && std::binary_search(rows[p].begin(), rows[p].end(), I);
        for (int c = 0; c < 2; ++c) {
          if (o.prev[r]) { o.wprev[r][c] = wslot(I, 2 * p + c); o.gprev[r][c] = slot3(I, 2 * p + c); }
          o.exists[r][c] = 1;
          o.amap[r][c] = S3.a_map[slot3(I, t0 + c)];
        }
      }
    }
    if (q == np) {               // only the forward-substitution block: its rows of panel np - 1 are solved, nothing is updated
      V3Own& o = S3.own[(size_t)q * kV3Warps + 1 + n_blocks];
      o.chunk_step[0] = 0; o.chunk_n[0] = 0; o.chunk_dest[0] = -1; o.chunk_kind[0] = 1;
      o.n_chunks = 1;
      continue;
    }
    // --- early-update steps: sources K < 2 p (every column before panel p)
    S3.pan[q].step0 = (int32_t)(S3.steps.size() / 4);
    std::vector<std::vector<int32_t>> bsteps(n_blocks + 2);     // [0 .. n_blocks - 1] tile blocks, [n_blocks] rhs, [n_blocks + 1] diagonal
    for (int K = 0; K < 2 * p; ++K) {
      const int32_t b0 = wslot(t0, K), b1 = wslot(t1, K);
      if (b0 < 0 && b1 < 0) continue;
      const int32_t B0 = b0 < 0 ? Z : b0, B1 = b1 < 0 ? Z : b1;
      for (int b = 0; b < n_blocks; ++b) {
        const int Ia = R[2 * b], Ib = (2 * b + 1 < (int)R.size()) ? R[2 * b + 1] : -1;
        const int32_t a0 = wslot(Ia, K), a1 = Ib >= 0 ? wslot(Ib, K) : -1;
        if (a0 < 0 && a1 < 0) continue;
        bsteps[b].insert(bsteps[b].end(), {a0 < 0 ? Z : a0, a1 < 0 ? Z : a1, B0, B1});
      }
      bsteps[n_blocks].insert(bsteps[n_blocks].end(), {K, 0, B0, B1});
      bsteps[n_blocks + 1].insert(bsteps[n_blocks + 1].end(), {B0, B1, B0, B1});
    }
    // --- chunks
    double weight = 0;
    for (int b = 0; b <= n_blocks + 1; ++b) weight += (b == n_blocks ? 0.5 : 1.0) * (double)(bsteps[b].size() / 4);
    const int lmax = std::max(4, (int)(weight / (kV3Warps - 1)) + 2);
    std::vector<double> load(kV3Warps, 0.0);
    std::vector<Chunk> helpers;
    std::vector<std::vector<int>> folds(n_blocks + 2);
    for (int b = 0; b <= n_blocks + 1; ++b) {
      const int n = (int)bsteps[b].size() / 4;
      const int32_t step0 = (int32_t)(S3.steps.size() / 4);
      S3.steps.insert(S3.steps.end(), bsteps[b].begin(), bsteps[b].end());
      S3.flops += (int64_t)n * (b == n_blocks ? 4 : 8) * 512;
      const int kind = b == n_blocks ? 1 : 0;
      const int n_ch = std::min((n + lmax - 1) / lmax, kV3MaxFold + (b <= n_blocks ? 1 : 0));
      for (int k = 0; k < n_ch; ++k) {
        const int s0 = (int)((int64_t)n * k / n_ch), s1 = (int)((int64_t)n * (k + 1) / n_ch);
        if (k == 0 && b <= n_blocks) {                         // the owner's own chunk
          V3Own& o = S3.own[(size_t)q * kV3Warps + 1 + b];
          o.chunk_step[0] = step0 + s0; o.chunk_n[0] = s1 - s0; o.chunk_dest[0] = -1; o.chunk_kind[0] = kind;
          o.n_chunks = 1;
          load[1 + b] += (kind ? 0.5 : 1.0) * (s1 - s0);
        } else {
          helpers.push_back({b, step0 + s0, s1 - s0, kind});
        }
      }
      if (n_ch == 0 && b <= n_blocks) {                         // nothing to subtract: the owner still initialises its block
        V3Own& o = S3.own[(size_t)q * kV3Warps + 1 + b];
        o.chunk_step[0] = 0; o.chunk_n[0] = 0; o.chunk_dest[0] = -1; o.chunk_kind[0] = kind;
        o.n_chunks = 1;
      }
    }
    S3.pan[q].n_steps = (int32_t)(S3.steps.size() / 4) - S3.pan[q].step0;
    S3.max_steps = std::max(S3.max_steps, S3.pan[q].n_steps);
    // longest helper chunks first, each to the least loaded update warp
    std::stable_sort(helpers.begin(), helpers.end(), [](const Chunk& a, const Chunk& b) { return a.n * (a.kind ? 1 : 2) > b.n * (b.kind ? 1 : 2); });
    int n_part = 0;
    for (const Chunk& c : helpers) {
      int best = -1;
      for (int w = 1; w < kV3Warps; ++w) {
        if (S3.own[(size_t)q * kV3Warps + w].n_chunks >= kV3MaxChunks) continue;
        if (best < 0 || load[w] < load[best]) best = w;
      }
      if (best < 0) return no("too many early-update chunks per warp");
      V3Own& o = S3.own[(size_t)q * kV3Warps + best];
      if (o.n_chunks == 0) {          // a warp without a block of its own: keep slot 0 for "no own block"
        o.chunk_step[0] = 0; o.chunk_n[0] = 0; o.chunk_dest[0] = -2; o.chunk_kind[0] = 0;
        o.n_chunks = 1;
      }
      const int k = o.n_chunks++;
      o.chunk_step[k] = c.step0; o.chunk_n[k] = c.n; o.chunk_dest[k] = n_part; o.chunk_kind[k] = c.kind;
      load[best] += (c.kind ? 0.5 : 1.0) * c.n;
      folds[c.block].push_back(n_part++);
    }
    S3.n_partial = std::max(S3.n_partial, n_part);
    for (int b = 0; b <= n_blocks + 1; ++b) {
      if ((int)folds[b].size() > kV3MaxFold) return no("a target block is split into too many chunks");
      int32_t* dst = (b == n_blocks + 1) ? S3.pan[q].fold : S3.own[(size_t)q * kV3Warps + 1 + b].fold;
      for (size_t k = 0; k < folds[b].size(); ++k) dst[k] = folds[b][k];
    }
  }
  // solves / factorisations: ~6 DMMAs per off-diagonal row of a panel + the diagonal blocks
  S3.flops += (n_tiles + (int64_t)np * 8) * 1024;
  S3.ok = true;
  return LRBMS_OK;
}

// ---- C access for tests (tests/test_symbolic3_emulator.py executes the tables in NumPy)
struct lrbms_symbolic_s;   // = lrbms_symbolic (opaque in the header)
extern "C" {

int lrbms_symbolic3_create(lrbms_symbolic_t s, lrbms_symbolic3_t* out) {
  if (!s || !out) return LRBMS_ERR_INVALID;
  lrbms_symbolic3* S3 = new lrbms_symbolic3();
  const int rc = lrbms_symbolic3_build(*S3, *s);
  if (rc) { delete S3; return rc; }
  *out = S3;
  return LRBMS_OK;
}

int lrbms_symbolic3_destroy(lrbms_symbolic3_t s) {
  delete s;
  return LRBMS_OK;
}

int lrbms_symbolic3_info(lrbms_symbolic3_t s, int32_t what, int64_t* out) {
  if (!s || !out) return LRBMS_ERR_INVALID;
  switch (what) {
    case 0: *out = s->ok ? 1 : 0; break;
    case 1: *out = s->np; break;
    case 2: *out = s->n_win_slots; break;
    case 3: *out = s->acc_rows; break;
    case 4: *out = s->n_partial; break;
    case 5: *out = s->n_tiles(); break;
    case 6: *out = s->flops; break;
    case 7: *out = (int64_t)(sizeof(V3Own) / 4); break;
    case 8: *out = (int64_t)(sizeof(V3Panel) / 4); break;
    case 9: *out = (int64_t)(s->steps.size() / 4); break;
    case 10: *out = (int64_t)s->why.size(); break;
    case 11: *out = s->max_steps; break;
    default: return LRBMS_ERR_INVALID;
  }
  return LRBMS_OK;
}

int64_t lrbms_symbolic3_get(lrbms_symbolic3_t s, int32_t which, int32_t* out, int64_t cap) {
  if (!s || !out) return LRBMS_ERR_INVALID;
  const int32_t* src = nullptr;
  int64_t n = 0;
  switch (which) {
    case 0: src = s->col_ptr.data(); n = (int64_t)s->col_ptr.size(); break;
    case 1: src = s->row_idx.data(); n = (int64_t)s->row_idx.size(); break;
    case 2: src = s->a_map.data(); n = (int64_t)s->a_map.size(); break;
    case 3: src = s->win_slot.data(); n = (int64_t)s->win_slot.size(); break;
    case 4: src = reinterpret_cast<const int32_t*>(s->own.data()); n = (int64_t)(s->own.size() * sizeof(V3Own) / 4); break;
    case 5: src = reinterpret_cast<const int32_t*>(s->pan.data()); n = (int64_t)(s->pan.size() * sizeof(V3Panel) / 4); break;
    case 6: src = s->steps.data(); n = (int64_t)s->steps.size(); break;
    case 7: {                                   // the reason the schedule does not apply, one character per int32
      n = std::min<int64_t>((int64_t)s->why.size(), cap);
      for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)s->why[i];
      return n;
    }
    default: return LRBMS_ERR_INVALID;
  }
  n = std::min<int64_t>(n, cap);
  std::copy(src, src + n, out);
  return n;
}

}  // extern "C"
