"""Seeded *synthetic* operator sets with the structure ``discretize()`` emits, in 2D or 3D.

DUNE-assembled operators cannot be produced here, and the host-side SWIPDG assembler of
:mod:`pylrbms_b200.swipdg_fixture` is 2D only.  For the 3D configuration of BASELINE.json (config C4: 8x8x8
subdomains, 6 face neighbours, local basis size 40) this module generates an operator set that has the *layout* of
SURVEY.md Appendix B -- the same dictionary of blocks, the same sparsity structure class, the same symmetry and
definiteness where the hot path relies on it -- from global random sparse matrices on a structured cell grid:

* every cell carries ``e`` DG dofs (4 = P1 on tetrahedra) and ``rt_per_cell`` flux dofs and is coupled to itself and to
  four face neighbours (x-1, x+1, and one neighbour each in y and z, chosen by parity so that the relation is
  symmetric) -> ``5 e`` non-zeros per row, the "20 nnz/row" of SURVEY.md section 8d row C4;
* global matrices with that cell pattern are sliced into subdomain blocks, so coupling blocks ``A_q[i, j]`` are
  non-zero on interface rows only (reference ``discretize...:560-561``) and the Oswald / flux-reconstruction
  components that leave a subdomain touch interface cells only;
* ``A(mu) = sum_q theta_q(mu) A_q`` is symmetric positive definite over the parameter range (strictly diagonally
  dominant), the local products are SPD, ``aa[q][q'] = aa[q'][q]^T``, ``bb`` is symmetric.

The values mean nothing physically; what the parity tests check is that the CUDA path reproduces the oracle on the
same operators, block by block, including 3D neighbourhoods with up to seven members.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import scipy.sparse as sp

from .swipdg_fixture import BlockSwipdgData, _csr32, os2015_problem


def _cell_layout(grid_shape, cps):
    """Global cell ids (subdomain-major) and the directed cell-neighbour pairs of the structured cell grid."""
    dim = len(grid_shape)
    gs = tuple(grid_shape) + (1,) * (3 - dim)
    h = tuple(cps) + (1,) * (3 - dim)
    G = tuple(gs[d] * h[d] for d in range(3))
    X, Y, Z = np.meshgrid(np.arange(G[0]), np.arange(G[1]), np.arange(G[2]), indexing='ij')
    ncs = h[0] * h[1] * h[2]

    def gid(x, y, z):
        sub = (x // h[0]) + gs[0] * ((y // h[1]) + gs[1] * (z // h[2]))
        loc = (x % h[0]) + h[0] * ((y % h[1]) + h[1] * (z % h[2]))
        return sub * ncs + loc

    me = gid(X, Y, Z)
    par = (X + Y + Z) % 2
    s = np.where(par == 0, 1, -1)
    pairs_a, pairs_b = [me.ravel()], [me.ravel()]
    candidates = [(X - 1, Y, Z), (X + 1, Y, Z), (X, Y + s, Z)] + ([(X, Y, Z + s)] if dim == 3 else [])
    for (nx, ny, nz) in candidates:
        ok = (nx >= 0) & (nx < G[0]) & (ny >= 0) & (ny < G[1]) & (nz >= 0) & (nz < G[2])
        pairs_a.append(me[ok])
        pairs_b.append(gid(nx[ok], ny[ok], nz[ok]))
    a, b = np.concatenate(pairs_a), np.concatenate(pairs_b)
    coords = np.zeros((me.size, 3))
    coords[me.ravel()] = np.stack([(X.ravel() + 0.5) / G[0], (Y.ravel() + 0.5) / G[1], (Z.ravel() + 0.5) / G[2]], axis=1)
    return gs, h, ncs, a, b, coords


def _block_matrix(rng, a, b, er, ec, n_cells, scale=1.0):
    """Random sparse matrix with an ``er x ec`` block for every directed cell pair ``(a, b)``."""
    P = a.size
    rows = (a[:, None, None] * er + np.arange(er)[None, :, None]) + np.zeros((1, 1, ec), dtype=np.int64)
    cols = (b[:, None, None] * ec + np.arange(ec)[None, None, :]) + np.zeros((1, er, 1), dtype=np.int64)
    vals = scale * rng.uniform(-1.0, 1.0, size=(P, er, ec))
    M = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(n_cells * er, n_cells * ec)).tocsr()
    M.sum_duplicates()
    return M


def _spd(rng, a, b, e, n_cells, shift=1.0):
    """Symmetric, strictly diagonally dominant (hence SPD) matrix with the cell pattern."""
    M = _block_matrix(rng, a, b, e, e, n_cells)
    M = (0.5 * (M + M.T)).tocsr()
    d = np.asarray(abs(M).sum(axis=1)).ravel() - np.abs(M.diagonal())
    M.setdiag(d + shift + rng.uniform(0.0, 0.5, size=d.size))
    M.sort_indices()
    return M.tocsr()


def synthetic_block_operators(grid_shape: Sequence[int] = (2, 2, 2), cells_per_subdomain: Sequence[int] = (2, 2, 2),
                              dofs_per_cell: int = 4, rt_per_cell: int = 2, seed: int = 1004,
                              problem=None) -> BlockSwipdgData:
    """Synthetic operator set on ``grid_shape`` subdomains (2 or 3 entries) of ``cells_per_subdomain`` cells each.

    ``(8, 8, 8), (16, 16, 12)`` has the sizes of BASELINE.json config C4 (``n_i = 12 288``, 20 nnz/row, 6 face
    neighbours); tests use small grids."""
    problem = os2015_problem() if problem is None else problem
    rng = np.random.default_rng(seed)
    dim = len(grid_shape)
    gs, h, ncs, a, b, coords = _cell_layout(grid_shape, cells_per_subdomain)
    S = gs[0] * gs[1] * gs[2]
    n_cells = S * ncs
    e, rt = int(dofs_per_cell), int(rt_per_cell)
    n_loc, m_loc = e * ncs, rt * ncs
    Q = len(problem['coefficients'])
    assert Q == 2, 'the synthetic generator follows the two-term OS2015 parametrisation'

    # ---- subdomain graph (face neighbours)
    neighbors, neighborhoods, boundary = [], [], []
    for s in range(S):
        ix, iy, iz = s % gs[0], (s // gs[0]) % gs[1], s // (gs[0] * gs[1])
        nb = []
        for d_, (i_, n_, stride) in enumerate(((ix, gs[0], 1), (iy, gs[1], gs[0]), (iz, gs[2], gs[0] * gs[1]))):
            if i_ > 0: nb.append(s - stride)
            if i_ < n_ - 1: nb.append(s + stride)
        neighbors.append(sorted(nb))
        neighborhoods.append(sorted(nb + [s]))
        if any(i_ in (0, n_ - 1) for (i_, n_) in ((ix, gs[0]), (iy, gs[1]), (iz, gs[2]))[:dim]):
            boundary.append(s)

    def sub_rows(i, per):
        return slice(i * ncs * per, (i + 1) * ncs * per)

    # ---- lhs: A(mu) = P0 + (1 - 0.8 mu) P1 with P0, P1 SPD  ->  A_0 = P0 + P1, A_1 = -0.8 P1  (theta = (1, mu), mu <= 1)
    P0, P1 = _spd(rng, a, b, e, n_cells), _spd(rng, a, b, e, n_cells)
    A = [(P0 + P1).tocsr(), (-0.8 * P1).tocsr()]
    lhs = []
    for q in range(Q):
        blocks = {}
        for i in range(S):
            Ri = A[q][sub_rows(i, e), :].tocsc()
            for j in neighborhoods[i]:
                blocks[(i, j)] = _csr32(Ri[:, sub_rows(j, e)].tocsr())
        lhs.append(blocks)

    # ---- local operators: restrict the cell pattern to one subdomain
    inner = (a // ncs) == (b // ncs)
    la, lb = (a[inner] % ncs), (b[inner] % ncs)
    own = (a[inner] // ncs)

    def local_pairs(i):
        sel = own == i
        return la[sel], lb[sel]

    mu_bar = float(np.asarray(problem['mu_bar']['diffusion']).ravel()[0])
    l2, energy, elliptic, div, bb, ab, aa, rhs = [], [], [], [], [], [[] for _ in range(Q)], \
        [[[] for _ in range(Q)] for _ in range(Q)], []
    for i in range(S):
        pa, pb = local_pairs(i)
        diag = pa == pb
        l2.append(_csr32(_spd(rng, pa[diag], pb[diag], e, ncs, shift=0.5)))        # block-diagonal element mass matrices
        energy.append(_csr32((lhs[0][(i, i)] + mu_bar * lhs[1][(i, i)]).tocsr()))   # the local block of A(mu_bar)
        elliptic.append(_csr32(_spd(rng, pa, pb, e, ncs)))
        div.append(_csr32(_block_matrix(rng, pa, pb, e, rt, ncs)[:n_loc, :m_loc]))
        Bm = _block_matrix(rng, pa, pb, rt, rt, ncs)
        bb.append(_csr32((0.5 * (Bm + Bm.T)).tocsr()))
        for q in range(Q):
            ab[q].append(_csr32(_block_matrix(rng, pa, pb, e, rt, ncs)))
            for q2 in range(q, Q):
                M = _block_matrix(rng, pa, pb, e, e, ncs)
                if q2 == q:
                    M = (0.5 * (M + M.T)).tocsr()
                aa[q][q2].append(_csr32(M))
        rhs.append(rng.standard_normal(n_loc))
    for q in range(Q):
        for q2 in range(q):
            aa[q][q2] = [_csr32(M.T.tocsr()) for M in aa[q2][q]]

    # ---- Oswald error / flux reconstruction: global matrices with the cell pattern, sliced per (source k, component i)
    OI = _block_matrix(rng, a, b, e, e, n_cells, scale=0.3)
    FR = [_block_matrix(rng, a, b, rt, e, n_cells, scale=0.5) for _ in range(Q)]
    oi, fr = {}, [dict() for _ in range(Q)]
    for i in range(S):
        Oi = OI[sub_rows(i, e), :].tocsc()
        Fi = [FR[q][sub_rows(i, rt), :].tocsc() for q in range(Q)]
        for k in neighborhoods[i]:
            oi[(k, i)] = _csr32(Oi[:, sub_rows(k, e)].tocsr())
            for q in range(Q):
                fr[q][(k, i)] = _csr32(Fi[q][:, sub_rows(k, e)].tocsr())

    # ---- scalars, shape functions (1, x, y, xy at the dofs of a cell: the cell centre plus a small per-dof offset)
    dof_xyz = np.repeat(coords, e, axis=0) + 1e-3 * np.tile(np.arange(e)[:, None], (n_cells, 3)) * np.array([1.0, -1.0, 0.5])
    dof_coords = [np.ascontiguousarray(dof_xyz[sub_rows(i, e), :dim]) for i in range(S)]
    shape_functions = []
    for i in range(S):
        x, y = dof_coords[i][:, 0], dof_coords[i][:, 1]
        shape_functions.append(np.stack([np.ones_like(x), x, y, x * y]))
    n = np.full(S, n_loc, dtype=np.int64)
    m = np.full(S, m_loc, dtype=np.int64)
    return BlockSwipdgData(
        num_subdomains=S, grid_shape=tuple(grid_shape), neighborhoods=neighborhoods, neighbors=neighbors,
        boundary_subdomains=boundary, n=n, m=m, coefficients=list(problem['coefficients']),
        parameter_type=dict(problem['parameter_type']), parameter_range=tuple(problem['parameter_range']),
        mu_bar=problem['mu_bar'], mu_hat=problem['mu_hat'], lhs=lhs, rhs=rhs, l2=l2, energy=energy, elliptic=elliptic,
        div=div, bb=bb, ab=ab, aa=aa, oi=oi, fr=fr, min_diffusion_evs=rng.uniform(0.5, 1.5, size=S),
        subdomain_diameters=np.full(S, float(np.sqrt(sum((1.0 / g) ** 2 for g in gs[:dim])))),
        local_eta_rf_squared=rng.uniform(0.5, 2.0, size=S), shape_functions=shape_functions, dof_coords=dof_coords,
        meta=dict(problem='synthetic-{}d'.format(dim), cells_per_subdomain=tuple(cells_per_subdomain), dim=dim, seed=seed,
                  synthetic=True))


def make_random_local_bases(data: BlockSwipdgData, basis_size, seed=0):
    """``[V_i]`` of shape ``(N_i, n_i)``, orthonormal in the local energy product: the four DG shape functions first
    (reference ``discretize...:187-200``), then seeded random vectors (``basis_size``: int or per-subdomain sequence)."""
    rng = np.random.default_rng(seed)
    S = data.num_subdomains
    sizes = [int(basis_size)] * S if np.isscalar(basis_size) else [int(b) for b in basis_size]
    bases = []
    for i in range(S):
        N, n_i = sizes[i], int(data.n[i])
        assert N <= n_i
        V = np.empty((N, n_i))
        k = min(4, N)
        V[:k] = data.shape_functions[i][:k]
        V[k:] = rng.standard_normal((N - k, n_i))
        E = data.energy[i]
        for _ in range(2):                                # two rounds of Cholesky-based orthonormalisation
            G = V @ (E @ V.T)
            L = np.linalg.cholesky(0.5 * (G + G.T))
            V = np.linalg.solve(L, V)
        bases.append(np.ascontiguousarray(V))
    return bases
