// Host-only symbolic phase of the mu-batched reduced solve (no CUDA calls in this file).
//
// The reduced LRBMS operator is block sparse: one N_i x N_j block per subdomain pair that shares a face
// (reference discretize_elliptic_block_swipdg.py:475-507).  The reference unblocks it into one dense matrix and
// calls numpy.linalg.solve.  Here the scalar index space is cut into 8x8 tiles -- the
// accumulator shape of DMMA.8x8x4 -- irrespective of subdomain boundaries, a symbolic Cholesky is run on the
// tile graph, and the numeric kernel gets a left-looking schedule: for every tile (I, J) of L the list of
// (L_IK, L_JK) pairs, K < J, whose product it has to subtract.  Ragged N_i need no padding except at the very
// end of the matrix.
#include "symbolic.h"

#include <algorithm>
#include <set>
#include <string>

#include "../../include/lrbms_sm100.h"

int lrbms_symbolic_build(lrbms_symbolic& S, int32_t n_sub, const int32_t* sizes, int32_t n_blocks, const int32_t* bi,
                         const int32_t* bj, std::string* err) {
  auto fail = [&](const char* m) { if (err) *err = m; return (int)LRBMS_ERR_INVALID; };
  if (n_sub <= 0 || !sizes || n_blocks < 0 || (n_blocks && (!bi || !bj))) return fail("symbolic: bad arguments");
  S.n_sub = n_sub;
  S.sizes.assign(sizes, sizes + n_sub);
  S.offsets.assign(n_sub + 1, 0);
  for (int i = 0; i < n_sub; ++i) {
    if (sizes[i] < 0) return fail("symbolic: negative basis size");
    S.offsets[i + 1] = S.offsets[i] + sizes[i];
  }
  S.n_red = S.offsets[n_sub];
  if (S.n_red == 0) return fail("symbolic: empty reduced space");
  S.n_pad = (S.n_red + 7) & ~7;
  S.ntc = S.n_pad / 8;
  S.block_i.assign(bi, bi + n_blocks);
  S.block_j.assign(bj, bj + n_blocks);

  // ---- tile pattern of the lower triangle of A
  std::vector<std::vector<int32_t>> cols(S.ntc);   // cols[J] = tile rows I >= J with A_IJ != 0
  std::vector<char> has_diag(n_sub, 0);
  for (int b = 0; b < n_blocks; ++b) {
    const int i = bi[b], j = bj[b];
    if (i < 0 || i >= n_sub || j < 0 || j >= n_sub) return fail("symbolic: block index out of range");
    if (i < j) continue;   // symmetric operator: the upper blocks are transposes of the lower ones
    if (i == j) has_diag[i] = 1;
    if (sizes[i] == 0 || sizes[j] == 0) continue;
    const int I0 = S.offsets[i] / 8, I1 = (S.offsets[i + 1] - 1) / 8;
    const int J0 = S.offsets[j] / 8, J1 = (S.offsets[j + 1] - 1) / 8;
    for (int J = J0; J <= J1; ++J)
      for (int I = std::max(I0, J); I <= I1; ++I) cols[J].push_back(I);
  }
  for (int i = 0; i < n_sub; ++i)
    if (sizes[i] > 0 && !has_diag[i]) return fail("symbolic: every subdomain needs its diagonal block");
  for (int J = 0; J < S.ntc; ++J) {
    cols[J].push_back(J);
    std::sort(cols[J].begin(), cols[J].end());
    cols[J].erase(std::unique(cols[J].begin(), cols[J].end()), cols[J].end());
  }
  std::vector<std::vector<int32_t>> a_cols = cols;

  // ---- symbolic Cholesky on the tile graph (elimination-tree merge)
  std::vector<std::vector<int32_t>> children(S.ntc);
  for (int J = 0; J < S.ntc; ++J) {
    std::vector<int32_t>& sj = cols[J];
    for (int K : children[J]) {
      // merge struct(K) \ {K} into struct(J)
      std::vector<int32_t> merged;
      const std::vector<int32_t>& sk = cols[K];
      merged.reserve(sj.size() + sk.size());
      std::set_union(sj.begin(), sj.end(), sk.begin() + 1, sk.end(), std::back_inserter(merged));
      sj.swap(merged);
    }
    // everything merged in is >= J by construction (parent = first off-diagonal row)
    if (sj.size() > 1) children[sj[1]].push_back(J);
  }
  S.col_ptr.assign(S.ntc + 1, 0);
  for (int J = 0; J < S.ntc; ++J) S.col_ptr[J + 1] = S.col_ptr[J] + (int32_t)cols[J].size();
  S.row_idx.clear();
  S.row_idx.reserve(S.col_ptr[S.ntc]);
  S.max_targets = 0;
  for (int J = 0; J < S.ntc; ++J) {
    S.row_idx.insert(S.row_idx.end(), cols[J].begin(), cols[J].end());
    S.max_targets = std::max<int32_t>(S.max_targets, (int32_t)cols[J].size() + 1);
  }
  const int64_t n_tiles = S.n_tiles();

  // ---- operator tiles
  S.a_map.assign(n_tiles, -1);
  S.n_a_tiles = 0;
  for (int J = 0; J < S.ntc; ++J) {
    size_t pa = 0;
    for (int32_t p = S.col_ptr[J]; p < S.col_ptr[J + 1]; ++p) {
      while (pa < a_cols[J].size() && a_cols[J][pa] < S.row_idx[p]) ++pa;
      if (pa < a_cols[J].size() && a_cols[J][pa] == S.row_idx[p]) S.a_map[p] = S.n_a_tiles++;
    }
  }

  // ---- left-looking update pairs.  Right-looking enumeration: column K contributes L_IK L_JK^T to every target
  //      (I, J) with J <= I both in struct(K) \ {K}; visiting K ascending yields each target's list sorted by K.
  const int64_t n_targets = n_tiles + S.ntc;
  std::vector<int64_t> count(n_targets, 0);
  // slot lookup inside a column: position by binary search
  auto slot_of = [&](int I, int J) -> int32_t {
    const int32_t* b = S.row_idx.data() + S.col_ptr[J];
    const int32_t* e = S.row_idx.data() + S.col_ptr[J + 1];
    const int32_t* it = std::lower_bound(b, e, I);
    return (int32_t)(it - S.row_idx.data());
  };
  int64_t total = 0;
  for (int K = 0; K < S.ntc; ++K) {
    const int32_t b = S.col_ptr[K] + 1, e = S.col_ptr[K + 1];
    for (int32_t pj = b; pj < e; ++pj) {
      const int J = S.row_idx[pj];
      for (int32_t pi = pj; pi < e; ++pi) { ++count[slot_of(S.row_idx[pi], J)]; ++total; }
      ++count[n_tiles + J]; ++total;     // forward-solve row: y_J -= y_K L_JK^T
    }
  }
  if (total > (int64_t)1 << 30) return fail("symbolic: reduced system too large for the CTA-per-parameter schedule");
  S.pair_ptr.assign(n_targets + 1, 0);
  for (int64_t tgt = 0; tgt < n_targets; ++tgt) S.pair_ptr[tgt + 1] = S.pair_ptr[tgt] + (int32_t)count[tgt];
  S.pair_a.assign(total, 0);
  S.pair_b.assign(total, 0);
  std::vector<int32_t> fill(S.pair_ptr.begin(), S.pair_ptr.end() - 1);
  for (int K = 0; K < S.ntc; ++K) {
    const int32_t b = S.col_ptr[K] + 1, e = S.col_ptr[K + 1];
    for (int32_t pj = b; pj < e; ++pj) {
      const int J = S.row_idx[pj];
      for (int32_t pi = pj; pi < e; ++pi) {
        const int32_t tgt = slot_of(S.row_idx[pi], J);
        S.pair_a[fill[tgt]] = pi;
        S.pair_b[fill[tgt]] = pj;
        ++fill[tgt];
      }
      const int64_t tgt = n_tiles + J;
      S.pair_a[fill[tgt]] = (int32_t)(n_tiles + K);
      S.pair_b[fill[tgt]] = pj;
      ++fill[tgt];
    }
  }
  // flops per mu: 2 * 512 per pair on L targets (8x8x8 multiply-add), plus potrf / trsm ~ 2 * 512 per tile
  S.flops = 0;
  for (int64_t tgt = 0; tgt < n_tiles; ++tgt) S.flops += (int64_t)count[tgt] * 1024;
  S.flops += n_tiles * 1024;
  return LRBMS_OK;
}

extern "C" {

int lrbms_symbolic_create(int32_t n_sub, const int32_t* basis_sizes, int32_t n_blocks, const int32_t* block_i,
                          const int32_t* block_j, lrbms_symbolic_t* out) {
  if (!out) return LRBMS_ERR_INVALID;
  *out = nullptr;
  lrbms_symbolic* S = new lrbms_symbolic();
  std::string err;
  int rc = lrbms_symbolic_build(*S, n_sub, basis_sizes, n_blocks, block_i, block_j, &err);
  if (rc) { delete S; return rc; }
  *out = S;
  return LRBMS_OK;
}

int lrbms_symbolic_destroy(lrbms_symbolic_t s) {
  delete s;
  return LRBMS_OK;
}

int lrbms_symbolic_info(lrbms_symbolic_t s, int32_t what, int64_t* out) {
  if (!s || !out) return LRBMS_ERR_INVALID;
  switch (what) {
    case 0: *out = s->n_red; break;
    case 1: *out = s->n_pad; break;
    case 2: *out = s->ntc; break;
    case 3: *out = s->n_tiles(); break;
    case 4: *out = s->n_a_tiles; break;
    case 5: *out = s->n_pairs(); break;
    case 6: *out = s->flops; break;
    case 7: *out = s->max_targets; break;
    default: return LRBMS_ERR_INVALID;
  }
  return LRBMS_OK;
}

int64_t lrbms_symbolic_get(lrbms_symbolic_t s, int32_t which, int32_t* out, int64_t cap) {
  if (!s || !out) return LRBMS_ERR_INVALID;
  const std::vector<int32_t>* v = nullptr;
  switch (which) {
    case 0: v = &s->col_ptr; break;
    case 1: v = &s->row_idx; break;
    case 2: v = &s->pair_ptr; break;
    case 3: v = &s->pair_a; break;
    case 4: v = &s->pair_b; break;
    case 5: v = &s->a_map; break;
    default: return LRBMS_ERR_INVALID;
  }
  const int64_t n = std::min<int64_t>(cap, (int64_t)v->size());
  std::copy(v->begin(), v->begin() + n, out);
  return n;
}

}  // extern "C"
