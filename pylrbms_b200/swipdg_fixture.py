"""Structured-grid P1 (S)IPDG assembler that emits the operator set of pylrbms' ``discretize()``.

pylrbms builds its block operators with dune-gdt (reference
``python/dune/pylrbms/discretize_elliptic_block_swipdg.py:530-811``).  DUNE is not available here and the
assembly itself is out of the hot-path scope, but the hot path needs inputs with
the *structure* that function emits.  This module produces them with plain
NumPy/SciPy on the host:

* domain ``[-1, 1]^2`` split into ``sx * sy`` square subdomains, ``h * h`` cells each, two triangles per cell,
  vertex-nodal P1-DG (3 dofs per triangle) -- ``n_i = 6 h^2`` fine dofs per subdomain;
* per affine diffusion component ``lambda_q`` a block operator with diagonal blocks (volume + inner faces +
  Dirichlet boundary faces + own side of the subdomain interfaces) and coupling blocks that are non-zero only
  on interface-element rows (reference ``:475-507``);
* right-hand side, local L2 / energy / elliptic products, and the estimator operators: divergence,
  diffusive-flux ``aa / bb / ab`` products, an RT0 diffusive-flux reconstruction and the Oswald interpolation
  error, the latter two as *sparse matrices* per (source subdomain, neighbourhood component)
  (reference ``:72-176, 639-770``).

Everything is ``float64`` / ``int32`` host data (``scipy.sparse.csr_matrix`` + ``numpy``).  The result is
consumed by the CPU oracle (tests only) and by :mod:`pylrbms_b200.discretization`, which uploads it to HBM.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------------------------------------
#  problem definitions (data functions + parameter functionals as expression strings)
# ----------------------------------------------------------------------------------------------------------

def _cos_bump(x, y):
    return np.cos(0.5 * np.pi * x) * np.cos(0.5 * np.pi * y)


def os2015_problem(mu_bar=1.0, mu_hat=1.0):
    """The OS2015 academic multiscale example (reference ``OS2015_academic_problem.py:19-67``)."""
    return dict(
        name='OS2015',
        lambdas=[lambda x, y: 1.0 + _cos_bump(x, y), lambda x, y: -1.0 * _cos_bump(x, y)],
        coefficients=['1.', 'diffusion'],                      # OS2015_academic_problem.py:43-44
        parameter_type={'diffusion': (1,)},
        f=lambda x, y: 0.5 * np.pi * np.pi * _cos_bump(x, y),
        lambda_bar=lambda x, y: 1.0 + (1.0 - mu_bar) * _cos_bump(x, y),
        lambda_hat=lambda x, y: 1.0 + (1.0 - mu_hat) * _cos_bump(x, y),
        mu_bar={'diffusion': np.array([mu_bar])},
        mu_hat={'diffusion': np.array([mu_hat])},
        parameter_range=(min(0.1, mu_bar, mu_hat), max(1.0, mu_bar, mu_hat)),
    )


def spe10_like_problem(seed=1003, contrast=1e6, mu_bar=0.5, mu_hat=0.5):
    """High-contrast channelised field (BASELINE.json config 3).

    ``lambda(mu) = 1 * lambda_0 + (1 - mu) * lambda_1`` with ``lambda_0 = 1`` and ``lambda_1`` = exp of a smooth
    seeded Gaussian-like field scaled to the requested contrast.
    """
    rng = np.random.default_rng(seed)
    nmodes = 24
    kx = rng.integers(1, 9, nmodes)
    ky = rng.integers(1, 9, nmodes)
    ph = rng.uniform(0, 2 * np.pi, (nmodes, 2))
    amp = rng.normal(size=nmodes) / np.sqrt(nmodes)
    log_c = 0.5 * np.log(contrast)

    def field_(x, y):
        g = np.zeros_like(x, dtype=float)
        for a, p, q, (p0, p1) in zip(amp, kx, ky, ph):
            g = g + a * np.sin(np.pi * p * x + p0) * np.sin(np.pi * q * y + p1)
        g = np.clip(g * 2.0, -1.0, 1.0)
        return np.exp(log_c * g)

    return dict(
        name='SPE10-like',
        lambdas=[lambda x, y: np.ones_like(x, dtype=float), field_],
        coefficients=['1.', '1. - diffusion'],
        parameter_type={'diffusion': (1,)},
        f=lambda x, y: np.ones_like(x, dtype=float),
        lambda_bar=lambda x, y: 1.0 + (1.0 - mu_bar) * field_(x, y),
        lambda_hat=lambda x, y: 1.0 + (1.0 - mu_hat) * field_(x, y),
        mu_bar={'diffusion': np.array([mu_bar])},
        mu_hat={'diffusion': np.array([mu_hat])},
        parameter_range=(0.1, 0.9),
    )


# ----------------------------------------------------------------------------------------------------------
#  result container
# ----------------------------------------------------------------------------------------------------------

@dataclass
class BlockSwipdgData:
    """Host-side operator set with the layout of SURVEY.md Appendix B (all CSR float64 / int32 indices)."""
    num_subdomains: int
    grid_shape: Tuple[int, ...]
    neighborhoods: List[List[int]]                 # grid.neighborhood_of(i): sorted, contains i
    neighbors: List[List[int]]                     # grid.neighboring_subdomains(i): face neighbours
    boundary_subdomains: List[int]
    n: np.ndarray                                  # fine DG dofs per subdomain
    m: np.ndarray                                  # local RT0 dofs per subdomain
    coefficients: List[str]                        # theta_q expressions
    parameter_type: Dict[str, tuple]
    parameter_range: Tuple[float, float]
    mu_bar: dict
    mu_hat: dict
    lhs: List[Dict[Tuple[int, int], sp.csr_matrix]]    # [q][(i, j)]  -> n_i x n_j
    rhs: List[np.ndarray]                              # [i] -> (n_i,)    (single rhs term, coefficient 1)
    l2: List[sp.csr_matrix]                            # [i] -> n_i x n_i
    energy: List[sp.csr_matrix]                        # [i] local_energy_dg_product_i assembled at mu_bar
    elliptic: List[sp.csr_matrix]                      # [i] local elliptic product for lambda_bar
    div: List[sp.csr_matrix]                           # [i] n_i x m_i
    bb: List[sp.csr_matrix]                            # [i] m_i x m_i
    ab: List[List[sp.csr_matrix]]                      # [q][i] n_i x m_i
    aa: List[List[List[sp.csr_matrix]]]                # [q][q'][i] n_i x n_i
    oi: Dict[Tuple[int, int], sp.csr_matrix]           # (k, i) i in N(k): n_i x n_k  (component i of OI_k)
    fr: List[Dict[Tuple[int, int], sp.csr_matrix]]     # [q][(k, i)]: m_i x n_k (component i of FR_k^q)
    min_diffusion_evs: np.ndarray
    subdomain_diameters: np.ndarray
    local_eta_rf_squared: np.ndarray
    shape_functions: List[np.ndarray]                  # [i] -> (4, n_i): 1, x, y, xy nodal interpolants
    dof_coords: List[np.ndarray]                       # [i] -> (n_i, dim)
    meta: dict = field(default_factory=dict)

    @property
    def Q(self):
        return len(self.coefficients)


# ----------------------------------------------------------------------------------------------------------
#  mesh
# ----------------------------------------------------------------------------------------------------------

class _TriMesh:
    """Structured triangle mesh of [-1,1]^2, elements numbered subdomain-major."""

    def __init__(self, sx, sy, h):
        self.sx, self.sy, self.h = sx, sy, h
        NX, NY = sx * h, sy * h
        self.NX, self.NY = NX, NY
        dx, dy = 2.0 / NX, 2.0 / NY
        cx, cy = np.meshgrid(np.arange(NX), np.arange(NY), indexing='xy')   # (NY, NX)
        sub = (cy // h) * sx + (cx // h)
        loc = ((cy % h) * h + (cx % h)) * 2
        nel_sub = 2 * h * h
        self.nel_sub = nel_sub
        self.elem_of_cell = sub * nel_sub + loc          # index of T0; T1 = +1   (NY, NX)
        nel = NX * NY * 2
        self.nel = nel
        x0 = -1.0 + cx * dx
        y0 = -1.0 + cy * dy
        x1 = x0 + dx
        y1 = y0 + dy
        verts = np.zeros((nel, 3, 2))
        e0 = self.elem_of_cell.ravel()
        X0, Y0, X1, Y1 = x0.ravel(), y0.ravel(), x1.ravel(), y1.ravel()
        verts[e0, 0] = np.c_[X0, Y0]
        verts[e0, 1] = np.c_[X1, Y0]
        verts[e0, 2] = np.c_[X1, Y1]
        verts[e0 + 1, 0] = np.c_[X0, Y0]
        verts[e0 + 1, 1] = np.c_[X1, Y1]
        verts[e0 + 1, 2] = np.c_[X0, Y1]
        self.verts = verts
        self.sub_of_elem = np.arange(nel) // nel_sub
        p0, p1, p2 = verts[:, 0], verts[:, 1], verts[:, 2]
        det = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1])
        assert np.all(det > 0)
        self.area = 0.5 * det
        grads = np.zeros((nel, 3, 2))
        for k in range(3):
            a, b = verts[:, (k + 1) % 3], verts[:, (k + 2) % 3]
            grads[:, k, 0] = (a[:, 1] - b[:, 1]) / det
            grads[:, k, 1] = (b[:, 0] - a[:, 0]) / det
        self.grads = grads
        self.centroid = verts.mean(axis=1)
        # integer vertex ids (for the Oswald interpolation)
        vid = np.zeros((nel, 3), dtype=np.int64)
        ix0, iy0 = cx.ravel(), cy.ravel()
        def V(ix, iy):
            return iy * (NX + 1) + ix
        vid[e0, 0] = V(ix0, iy0); vid[e0, 1] = V(ix0 + 1, iy0); vid[e0, 2] = V(ix0 + 1, iy0 + 1)
        vid[e0 + 1, 0] = V(ix0, iy0); vid[e0 + 1, 1] = V(ix0 + 1, iy0 + 1); vid[e0 + 1, 2] = V(ix0, iy0 + 1)
        self.vid = vid
        vx, vy = np.meshgrid(np.arange(NX + 1), np.arange(NY + 1), indexing='xy')
        self.vertex_on_boundary = ((vx == 0) | (vx == NX) | (vy == 0) | (vy == NY)).ravel()
        self._build_faces()

    def _build_faces(self):
        NX, NY = self.NX, self.NY
        E = self.elem_of_cell
        plus, eplus, minus, eminus = [], [], [], []
        # diagonal faces: T0 edge 1 | T1 edge 2
        plus.append(E.ravel()); eplus.append(np.full(E.size, 1)); minus.append(E.ravel() + 1); eminus.append(np.full(E.size, 2))
        # vertical faces, interior: left cell T0 edge 0 | right cell T1 edge 1
        if NX > 1:
            L = E[:, :-1].ravel(); R = E[:, 1:].ravel() + 1
            plus.append(L); eplus.append(np.full(L.size, 0)); minus.append(R); eminus.append(np.full(L.size, 1))
        # horizontal faces, interior: lower cell T1 edge 0 | upper cell T0 edge 2
        if NY > 1:
            B = E[:-1, :].ravel() + 1; T = E[1:, :].ravel()
            plus.append(B); eplus.append(np.full(B.size, 0)); minus.append(T); eminus.append(np.full(B.size, 2))
        # boundary faces
        bl = E[:, 0].ravel() + 1       # left boundary: T1 edge 1
        br = E[:, -1].ravel()          # right boundary: T0 edge 0
        bb = E[0, :].ravel()           # bottom: T0 edge 2
        bt = E[-1, :].ravel() + 1      # top: T1 edge 0
        for els, ed in ((bl, 1), (br, 0), (bb, 2), (bt, 0)):
            plus.append(els); eplus.append(np.full(els.size, ed)); minus.append(np.full(els.size, -1)); eminus.append(np.full(els.size, -1))
        self.f_plus = np.concatenate(plus)
        self.f_eplus = np.concatenate(eplus)
        self.f_minus = np.concatenate(minus)
        self.f_eminus = np.concatenate(eminus)
        nf = self.f_plus.size
        self.nf = nf
        ar = np.arange(nf)
        # face endpoints as seen from T+ (local vertices (e+1)%3, (e+2)%3)
        self.f_np = np.stack([(self.f_eplus + 1) % 3, (self.f_eplus + 2) % 3], axis=1)       # (nf, 2)
        A = self.verts[self.f_plus, self.f_np[:, 0]]
        B = self.verts[self.f_plus, self.f_np[:, 1]]
        t = B - A
        self.f_len = np.linalg.norm(t, axis=1)
        # CCW element => outward normal of T+ on the edge A->B is (t_y, -t_x)/|t|
        self.f_normal = np.stack([t[:, 1], -t[:, 0]], axis=1) / self.f_len[:, None]
        self.f_mid = 0.5 * (A + B)
        interior = self.f_minus >= 0
        self.f_interior = interior
        # T- traverses the edge in the opposite direction: its (e'+1)%3 is B, (e'+2)%3 is A
        nm = np.zeros((nf, 2), dtype=np.int64)
        nm[interior, 0] = (self.f_eminus[interior] + 2) % 3     # matches A
        nm[interior, 1] = (self.f_eminus[interior] + 1) % 3     # matches B
        self.f_nm = nm
        fi = ar[interior]
        assert np.allclose(self.verts[self.f_minus[fi], nm[fi, 0]], A[fi])
        assert np.allclose(self.verts[self.f_minus[fi], nm[fi, 1]], B[fi])
        # outward check
        c = self.centroid[self.f_plus]
        assert np.all(np.einsum('ij,ij->i', self.f_mid - c, self.f_normal) > 0)
        # face id per (element, local edge)
        fe = np.full((self.nel, 3), -1, dtype=np.int64)
        fe[self.f_plus, self.f_eplus] = ar
        fe[self.f_minus[fi], self.f_eminus[fi]] = fi
        assert np.all(fe >= 0)
        self.face_of_elem = fe
        sg = np.zeros((self.nel, 3))
        sg[self.f_plus, self.f_eplus] = 1.0
        sg[self.f_minus[fi], self.f_eminus[fi]] = -1.0
        self.face_sign = sg      # +1 if the face normal is outward for the element


# ----------------------------------------------------------------------------------------------------------
#  assembly helpers
# ----------------------------------------------------------------------------------------------------------

_EDGE_MASS = np.array([[2.0, 1.0], [1.0, 2.0]]) / 6.0


def _coo(rows, cols, vals, shape):
    M = sp.coo_matrix((np.concatenate([np.ravel(v) for v in vals]),
                       (np.concatenate([np.ravel(r) for r in rows]), np.concatenate([np.ravel(c) for c in cols]))),
                      shape=shape).tocsr()
    M.sum_duplicates()
    M.sort_indices()
    return M


def _assemble_ipdg(mesh: _TriMesh, lam: np.ndarray, sigma0: float, face_mask=None, volume=True,
                   consistency=True, penalty=True, dirichlet_mask=None):
    """Global (all subdomains) symmetric IPDG matrix for an element-wise constant diffusion ``lam``.

    ``face_mask`` restricts the face terms; ``dirichlet_mask`` marks non-interior faces treated as Dirichlet
    (default: every boundary face).  The penalty is ``sigma0 * lam_F / |F|`` with the arithmetic face mean, so
    the whole form is linear in ``lam`` (affine decomposition stays exact).
    """
    nel, ndof = mesh.nel, 3 * mesh.nel
    rows, cols, vals = [], [], []
    if volume:
        K = (lam * mesh.area)[:, None, None] * np.einsum('eik,ejk->eij', mesh.grads, mesh.grads)
        d = 3 * np.arange(nel)[:, None] + np.arange(3)[None, :]
        rows.append(np.repeat(d[:, :, None], 3, axis=2)); cols.append(np.repeat(d[:, None, :], 3, axis=1)); vals.append(K)
    fsel = np.ones(mesh.nf, dtype=bool) if face_mask is None else face_mask
    # ---- interior faces
    fi = np.where(mesh.f_interior & fsel)[0]
    if fi.size:
        ep, em = mesh.f_plus[fi], mesh.f_minus[fi]
        n = mesh.f_normal[fi]
        L = mesh.f_len[fi]
        lp, lm = lam[ep], lam[em]
        dnp = np.einsum('fkd,fd->fk', mesh.grads[ep], n)      # d_n phi_k on T+   (nf, 3)
        dnm = np.einsum('fkd,fd->fk', mesh.grads[em], n)
        nodes_p = 3 * ep[:, None] + mesh.f_np[fi]             # (nf, 2) dofs on the face, ordered (A, B)
        nodes_m = 3 * em[:, None] + mesh.f_nm[fi]
        dofs_p = 3 * ep[:, None] + np.arange(3)[None, :]
        dofs_m = 3 * em[:, None] + np.arange(3)[None, :]
        if consistency:
            # C[v-node, u-dof] = -(+-1)_v * |F|/2 * 1/2 * lam_u * d_n phi_u ;  A += C + C^T
            for (nv, sv) in ((nodes_p, 1.0), (nodes_m, -1.0)):
                for (du, lu, dn) in ((dofs_p, lp, dnp), (dofs_m, lm, dnm)):
                    c = -sv * (0.5 * L * 0.5 * lu)[:, None] * dn                 # (nf, 3)
                    r = np.repeat(nv[:, :, None], 3, axis=2)                      # (nf, 2, 3)
                    cc = np.repeat(du[:, None, :], 2, axis=1)
                    v = np.repeat(c[:, None, :], 2, axis=1)
                    rows.append(r); cols.append(cc); vals.append(v)
                    rows.append(cc); cols.append(r); vals.append(v)
        if penalty:
            sig = sigma0 * 0.5 * (lp + lm) / L
            P = (sig * L)[:, None, None] * _EDGE_MASS[None]
            for (nv, sv) in ((nodes_p, 1.0), (nodes_m, -1.0)):
                for (nu, su) in ((nodes_p, 1.0), (nodes_m, -1.0)):
                    rows.append(np.repeat(nv[:, :, None], 2, axis=2)); cols.append(np.repeat(nu[:, None, :], 2, axis=1))
                    vals.append(sv * su * P)
    # ---- Dirichlet faces
    dm = (~mesh.f_interior) if dirichlet_mask is None else dirichlet_mask
    fb = np.where(dm & fsel)[0]
    if fb.size:
        ep = mesh.f_plus[fb]
        n = mesh.f_normal[fb]
        L = mesh.f_len[fb]
        lp = lam[ep]
        dnp = np.einsum('fkd,fd->fk', mesh.grads[ep], n)
        nodes_p = 3 * ep[:, None] + mesh.f_np[fb]
        dofs_p = 3 * ep[:, None] + np.arange(3)[None, :]
        if consistency:
            c = -(0.5 * L * lp)[:, None] * dnp
            r = np.repeat(nodes_p[:, :, None], 3, axis=2)
            cc = np.repeat(dofs_p[:, None, :], 2, axis=1)
            v = np.repeat(c[:, None, :], 2, axis=1)
            rows.append(r); cols.append(cc); vals.append(v)
            rows.append(cc); cols.append(r); vals.append(v)
        if penalty:
            sig = sigma0 * lp / L
            P = (sig * L)[:, None, None] * _EDGE_MASS[None]
            rows.append(np.repeat(nodes_p[:, :, None], 2, axis=2)); cols.append(np.repeat(nodes_p[:, None, :], 2, axis=1)); vals.append(P)
    return _coo(rows, cols, vals, (ndof, ndof))


def _csr32(M):
    M = sp.csr_matrix(M)
    M.sort_indices()
    M.indices = M.indices.astype(np.int32)
    M.indptr = M.indptr.astype(np.int32)
    M.data = np.ascontiguousarray(M.data, dtype=np.float64)
    return M


# ----------------------------------------------------------------------------------------------------------
#  main entry point
# ----------------------------------------------------------------------------------------------------------

def assemble_block_swipdg(num_subdomains: Sequence[int] = (4, 4), cells_per_subdomain: int = 16, problem=None,
                          sigma0: float = 8.0, keep_structural_zeros: bool = True) -> BlockSwipdgData:
    """Assemble the operator set of SURVEY.md Appendix B on a structured triangle grid.

    ``num_subdomains=(4, 4), cells_per_subdomain=16`` is config C1 (n_i = 1536), ``(8, 8), 32`` is C2/C5
    (n_i = 6144), ``(16, 16), 32`` is C3.
    """
    problem = os2015_problem() if problem is None else problem
    sx, sy = num_subdomains
    h = cells_per_subdomain
    mesh = _TriMesh(sx, sy, h)
    S = sx * sy
    nel_sub = mesh.nel_sub
    n_loc = 3 * nel_sub
    n = np.full(S, n_loc, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(n)])
    Q = len(problem['lambdas'])
    cx, cy = mesh.centroid[:, 0], mesh.centroid[:, 1]
    lam_q = [np.asarray(f(cx, cy), dtype=float) for f in problem['lambdas']]
    lam_bar = np.asarray(problem['lambda_bar'](cx, cy), dtype=float)
    lam_hat = np.asarray(problem['lambda_hat'](cx, cy), dtype=float)

    # ---- subdomain graph
    neighbors, neighborhoods = [], []
    for s in range(S):
        ix, iy = s % sx, s // sx
        nb = []
        if iy > 0: nb.append(s - sx)
        if ix > 0: nb.append(s - 1)
        if ix < sx - 1: nb.append(s + 1)
        if iy < sy - 1: nb.append(s + sx)
        neighbors.append(sorted(nb))
        neighborhoods.append(sorted(nb + [s]))
    boundary_subdomains = [s for s in range(S) if (s % sx in (0, sx - 1)) or (s // sx in (0, sy - 1))]

    def blk(M, i, j):
        return _csr32(M[off[i]:off[i + 1], :][:, off[j]:off[j + 1]])

    # ---- lhs: one block operator per affine component  (discretize...:475-507)
    lhs = []
    for q in range(Q):
        Aq = _assemble_ipdg(mesh, lam_q[q], sigma0)
        blocks = {}
        for i in range(S):
            blocks[(i, i)] = blk(Aq, i, i)
            for j in neighbors[i]:
                blocks[(i, j)] = blk(Aq, i, j)
        lhs.append(blocks)
    if keep_structural_zeros and Q > 1:
        # the reference allocates every component with the same pattern (local_patterns / coupling_patterns are
        # shared across lambda_funcs, discretize...:548-565); give all q the union pattern as well
        for key in lhs[0]:
            pat = sp.csr_matrix(sum((abs(lhs[q][key]) for q in range(Q))))
            for q in range(Q):
                lhs[q][key] = _union_pattern(lhs[q][key], pat)

    # ---- rhs (single term)  (discretize...:510-527)
    fvals_mid = []  # f at the three edge midpoints of each element
    V = mesh.verts
    mids = np.stack([0.5 * (V[:, 1] + V[:, 2]), 0.5 * (V[:, 2] + V[:, 0]), 0.5 * (V[:, 0] + V[:, 1])], axis=1)  # mid opposite vertex k
    fm = np.asarray(problem['f'](mids[..., 0], mids[..., 1]), dtype=float)       # (nel, 3)
    # phi_a at the midpoint opposite vertex k is 0 if a == k else 1/2
    W = 0.5 * (1.0 - np.eye(3))
    rhs_glob = (mesh.area[:, None] / 3.0) * (fm @ W.T)
    rhs_glob = rhs_glob.ravel()
    rhs = [np.ascontiguousarray(rhs_glob[off[i]:off[i + 1]]) for i in range(S)]
    f2 = (mesh.area / 3.0) * np.sum(fm ** 2, axis=1)
    local_eta_rf_squared = np.array([f2[i * nel_sub:(i + 1) * nel_sub].sum() for i in range(S)])

    # ---- local products
    el_dofs = 3 * np.arange(mesh.nel)[:, None] + np.arange(3)[None, :]
    Mloc = np.array([[2.0, 1.0, 1.0], [1.0, 2.0, 1.0], [1.0, 1.0, 2.0]]) / 12.0
    Mg = _coo([np.repeat(el_dofs[:, :, None], 3, axis=2)], [np.repeat(el_dofs[:, None, :], 3, axis=1)],
              [mesh.area[:, None, None] * Mloc[None]], (3 * mesh.nel,) * 2)
    l2 = [blk(Mg, i, i) for i in range(S)]
    Eg = _assemble_ipdg(mesh, lam_bar, sigma0, consistency=False, penalty=False)
    elliptic = [blk(Eg, i, i) for i in range(S)]
    # local energy product: elliptic + penalty with all-Dirichlet *subdomain* boundary, assembled at mu_bar
    # (discretize...:651-677).  Faces on subdomain interfaces are Dirichlet faces of both adjacent subdomains.
    theta_bar = evaluate_coefficients(problem['coefficients'], problem['mu_bar'])
    lam_mu_bar = sum(t * l for t, l in zip(theta_bar, lam_q))
    sub_p = mesh.sub_of_elem[mesh.f_plus]
    sub_m = np.where(mesh.f_interior, mesh.sub_of_elem[np.maximum(mesh.f_minus, 0)], -1)
    iface = mesh.f_interior & (sub_p != sub_m)
    inner = mesh.f_interior & ~iface
    Pen_in = _assemble_ipdg(mesh, lam_mu_bar, sigma0, face_mask=inner | ~mesh.f_interior, volume=True,
                            consistency=False, penalty=True)
    # interface faces: one-sided penalty for each side
    Pen_if = _interface_one_sided_penalty(mesh, lam_mu_bar, sigma0, iface)
    Eng = Pen_in + Pen_if
    energy = [blk(Eng, i, i) for i in range(S)]

    # ---- RT0 spaces per subdomain
    face_sub_lists = []
    m = np.zeros(S, dtype=np.int64)
    rt_index = []      # per subdomain: dict-like array global face -> local rt dof (or -1)
    for i in range(S):
        fe = mesh.face_of_elem[i * nel_sub:(i + 1) * nel_sub].ravel()
        faces = np.unique(fe)
        face_sub_lists.append(faces)
        m[i] = faces.size
        idx = np.full(mesh.nf, -1, dtype=np.int64)
        idx[faces] = np.arange(faces.size)
        rt_index.append(idx)

    # ---- estimator volume operators on each subdomain
    div, bb = [], []
    ab = [[None] * S for _ in range(Q)]
    aa = [[[None] * S for _ in range(Q)] for _ in range(Q)]
    GG = np.einsum('eik,ejk->eij', mesh.grads, mesh.grads)
    # RT0 basis on element T for local edge k: psi_k(x) = sign_k |F_k| / (2|T|) (x - p_k)
    for i in range(S):
        els = np.arange(i * nel_sub, (i + 1) * nel_sub)
        fe = mesh.face_of_elem[els]                      # (ne, 3) global faces
        rt = rt_index[i][fe]                             # (ne, 3) local rt dofs
        sg = mesh.face_sign[els]
        FL = mesh.f_len[fe]
        ar = mesh.area[els]
        ld = 3 * np.arange(nel_sub)[:, None] + np.arange(3)[None, :]          # local dg dofs
        # divergence: nodal DG coefficients of div psi_F (constant sign |F| / |T| on T), so that with the reference's
        # algebra  f_vec^T D t = (f, div t)  and  (D t)^T M (D t) = ||div t||^2   (reference estimators.py:72-76)
        dv = (sg * FL / ar[:, None])                                           # (ne, 3[F])
        div.append(_csr32(_coo([np.repeat(ld[:, :, None], 3, axis=2)], [np.repeat(rt[:, None, :], 3, axis=1)],
                               [np.repeat(dv[:, None, :], 3, axis=1)], (n_loc, m[i]))))
        # psi_k at the 3 edge midpoints: (ne, k, midpoint j, 2)
        coef = sg * FL / (2.0 * ar[:, None])                                   # (ne, 3)
        pm = mids[els][:, None, :, :] - V[els][:, :, None, :]                  # x_mid_j - p_k
        psi = coef[:, :, None, None] * pm                                      # (ne, 3k, 3j, 2)
        lh = lam_hat[els]
        Bl = np.einsum('ekjd,eljd->ekl', psi, psi) * (ar / 3.0 / lh)[:, None, None]
        bb.append(_csr32(_coo([np.repeat(rt[:, :, None], 3, axis=2)], [np.repeat(rt[:, None, :], 3, axis=1)], [Bl], (m[i], m[i]))))
        psi_int = coef[:, :, None] * (mesh.centroid[els][:, None, :] - V[els]) * ar[:, None, None]    # int_T psi_k (ne,3,2)
        for q in range(Q):
            w = lam_q[q][els] / lh
            Al = np.einsum('ead,ekd->eak', mesh.grads[els], psi_int) * w[:, None, None]
            ab[q][i] = _csr32(_coo([np.repeat(ld[:, :, None], 3, axis=2)], [np.repeat(rt[:, None, :], 3, axis=1)], [Al], (n_loc, m[i])))
            for q2 in range(Q):
                w2 = lam_q[q][els] * lam_q[q2][els] / lh * ar
                aa[q][q2][i] = _csr32(_coo([np.repeat(ld[:, :, None], 3, axis=2)], [np.repeat(ld[:, None, :], 3, axis=1)],
                                           [GG[els] * w2[:, None, None]], (n_loc, n_loc)))

    # ---- flux reconstruction FR^q: DG -> RT0 normal-flux dofs (global, then cut into (k, i) components)
    fr = []
    for q in range(Q):
        Fg = _assemble_flux_reconstruction(mesh, lam_q[q], sigma0)        # (nf, ndof)
        comp = {}
        for k in range(S):
            cols = Fg[:, off[k]:off[k + 1]].tocsr()
            for i in neighborhoods[k]:
                comp[(k, i)] = _csr32(cols[face_sub_lists[i], :])
        fr.append(comp)

    # ---- Oswald interpolation error  u - I_os(u)
    oi = {}
    for i in range(S):
        Oi = _assemble_oswald_error_rows(mesh, i, neighborhoods[i], off)     # rows of subdomain i, all columns
        for k in neighborhoods[i]:
            oi[(k, i)] = _csr32(Oi[:, off[k]:off[k + 1]])

    min_ev = np.array([lam_hat[i * nel_sub:(i + 1) * nel_sub].min() for i in range(S)])
    diam = np.full(S, np.sqrt((2.0 / sx) ** 2 + (2.0 / sy) ** 2))

    dofxy = mesh.verts.reshape(-1, 2)
    dof_coords = [np.ascontiguousarray(dofxy[off[i]:off[i + 1]]) for i in range(S)]
    shape_functions = []
    for i in range(S):
        x, y = dof_coords[i][:, 0], dof_coords[i][:, 1]
        shape_functions.append(np.stack([np.ones_like(x), x, y, x * y]))

    return BlockSwipdgData(
        num_subdomains=S, grid_shape=(sx, sy), neighborhoods=neighborhoods, neighbors=neighbors,
        boundary_subdomains=boundary_subdomains, n=n, m=m, coefficients=list(problem['coefficients']),
        parameter_type=dict(problem['parameter_type']), parameter_range=tuple(problem['parameter_range']),
        mu_bar=problem['mu_bar'], mu_hat=problem['mu_hat'], lhs=lhs, rhs=rhs, l2=l2, energy=energy,
        elliptic=elliptic, div=div, bb=bb, ab=ab, aa=aa, oi=oi, fr=fr, min_diffusion_evs=min_ev,
        subdomain_diameters=diam, local_eta_rf_squared=local_eta_rf_squared, shape_functions=shape_functions,
        dof_coords=dof_coords,
        meta=dict(problem=problem['name'], cells_per_subdomain=h, sigma0=sigma0, dim=2))


def _union_pattern(M, pat):
    """Return ``M`` stored on the pattern of ``pat`` (a CSR with explicit zeros)."""
    out = sp.csr_matrix(pat, copy=True)
    out.data = np.zeros_like(out.data)
    M = sp.csr_matrix(M); M.sort_indices()
    out.sort_indices()
    # positions of M's entries inside pat's pattern
    ncols = out.shape[1]
    key_out = np.repeat(np.arange(out.shape[0], dtype=np.int64), np.diff(out.indptr)) * ncols + out.indices
    key_m = np.repeat(np.arange(M.shape[0], dtype=np.int64), np.diff(M.indptr)) * ncols + M.indices
    pos = np.searchsorted(key_out, key_m)
    assert np.all(key_out[pos] == key_m)
    out.data[pos] = M.data
    return _csr32(out)


def _interface_one_sided_penalty(mesh, lam, sigma0, iface):
    fi = np.where(iface)[0]
    ndof = 3 * mesh.nel
    if fi.size == 0:
        return sp.csr_matrix((ndof, ndof))
    rows, cols, vals = [], [], []
    L = mesh.f_len[fi]
    for (el, nodes) in ((mesh.f_plus[fi], mesh.f_np[fi]), (mesh.f_minus[fi], mesh.f_nm[fi])):
        nd = 3 * el[:, None] + nodes
        P = (sigma0 * lam[el])[:, None, None] * _EDGE_MASS[None]        # sigma * L = sigma0 * lam
        rows.append(np.repeat(nd[:, :, None], 2, axis=2)); cols.append(np.repeat(nd[:, None, :], 2, axis=1)); vals.append(P)
    return _coo(rows, cols, vals, (ndof, ndof))


def _assemble_flux_reconstruction(mesh, lam, sigma0):
    """RT0 normal-flux dofs of the diffusive flux reconstruction:  t.n_F = -{lam d_n u} + sigma_F mean_F [u]."""
    ndof = 3 * mesh.nel
    rows, cols, vals = [], [], []
    fi = np.where(mesh.f_interior)[0]
    ep, em = mesh.f_plus[fi], mesh.f_minus[fi]
    n = mesh.f_normal[fi]
    L = mesh.f_len[fi]
    lp, lm = lam[ep], lam[em]
    sig = sigma0 * 0.5 * (lp + lm) / L
    for (el, lu, nodes, s) in ((ep, lp, mesh.f_np[fi], 1.0), (em, lm, mesh.f_nm[fi], -1.0)):
        dn = np.einsum('fkd,fd->fk', mesh.grads[el], n)
        dofs = 3 * el[:, None] + np.arange(3)[None, :]
        rows.append(np.repeat(fi[:, None], 3, axis=1)); cols.append(dofs); vals.append(-0.5 * lu[:, None] * dn)
        nd = 3 * el[:, None] + nodes
        rows.append(np.repeat(fi[:, None], 2, axis=1)); cols.append(nd); vals.append(np.repeat((s * 0.5 * sig)[:, None], 2, axis=1))
    fb = np.where(~mesh.f_interior)[0]
    ep = mesh.f_plus[fb]
    n = mesh.f_normal[fb]
    L = mesh.f_len[fb]
    lp = lam[ep]
    sig = sigma0 * lp / L
    dn = np.einsum('fkd,fd->fk', mesh.grads[ep], n)
    dofs = 3 * ep[:, None] + np.arange(3)[None, :]
    rows.append(np.repeat(fb[:, None], 3, axis=1)); cols.append(dofs); vals.append(-lp[:, None] * dn)
    nd = 3 * ep[:, None] + mesh.f_np[fb]
    rows.append(np.repeat(fb[:, None], 2, axis=1)); cols.append(nd); vals.append(np.repeat((0.5 * sig)[:, None], 2, axis=1))
    return _coo(rows, cols, vals, (mesh.nf, ndof))


def _assemble_oswald_error_rows(mesh, i, neighborhood, off):
    """Rows of subdomain ``i`` of ``I - Avg_i``.  ``Avg_i`` averages, at every mesh vertex, the DG values of the
    elements *inside the neighbourhood of i* (zero at Dirichlet boundary vertices) -- the reference interpolates on
    subdomain ``ii`` with only its neighbourhood's data available (``discretize...:90-119``)."""
    ndof = 3 * mesh.nel
    nel_sub = mesh.nel_sub
    els = np.concatenate([np.arange(k * nel_sub, (k + 1) * nel_sub) for k in neighborhood])
    dofs = (3 * els[:, None] + np.arange(3)[None, :]).ravel()
    vid = mesh.vid[els].ravel()
    nv = mesh.vertex_on_boundary.size
    cnt = np.bincount(vid, minlength=nv).astype(float)
    w = np.where(mesh.vertex_on_boundary | (cnt == 0), 0.0, 1.0 / np.maximum(cnt, 1.0))
    Pv = sp.csr_matrix((w[vid], (vid, dofs)), shape=(nv, ndof))
    rows_i = np.arange(off[i], off[i + 1])
    Inj = sp.csr_matrix((np.ones(rows_i.size), (np.arange(rows_i.size), mesh.vid.ravel()[rows_i])), shape=(rows_i.size, nv))
    Id = sp.csr_matrix((np.ones(rows_i.size), (np.arange(rows_i.size), rows_i)), shape=(rows_i.size, ndof))
    Oi = sp.csr_matrix(Id - Inj @ Pv)
    Oi.sort_indices()
    return Oi


# ----------------------------------------------------------------------------------------------------------
#  parameter functionals (host): expression strings over the parameter components, vectorised over a mu batch
# ----------------------------------------------------------------------------------------------------------

_SAFE = {'sin': np.sin, 'cos': np.cos, 'exp': np.exp, 'sqrt': np.sqrt, 'pi': np.pi, 'abs': np.abs,
         'min': np.minimum, 'max': np.maximum, 'log': np.log, 'tan': np.tan}


def evaluate_coefficients(expressions, mu):
    """Evaluate ``ExpressionParameterFunctional``-style strings (reference ``OS2015_academic_problem.py:43-44``,
    ``local_thermalblock_problem.py:50-51``).  ``mu`` maps component name -> array of shape ``(dim,)`` or
    ``(n_mu, dim)``; the result is a list of scalars or ``(n_mu,)`` arrays."""
    env = dict(_SAFE)
    batched = False
    for k, v in mu.items():
        v = np.asarray(v, dtype=float)
        if v.ndim == 2:
            batched = True
            env[k] = v.T if v.shape[1] > 1 else v[:, 0]
        else:
            env[k] = v if v.size > 1 else float(v.ravel()[0])
    out = []
    for e in expressions:
        if callable(e):
            val = e(mu)
        elif isinstance(e, (int, float)):
            val = float(e)
        else:
            val = eval(e, {'__builtins__': {}}, env)        # noqa: S307 - restricted namespace, trusted config strings
        if batched:
            n_mu = next(np.asarray(v).shape[0] for v in mu.values() if np.asarray(v).ndim == 2)
            val = np.broadcast_to(np.asarray(val, dtype=float), (n_mu,)).copy()
        else:
            val = float(val)
        out.append(val)
    return out


# ----------------------------------------------------------------------------------------------------------
#  local reduced bases (host): what the reference obtains from extend_basis(_local) with the local energy
#  products (reference online_adaptive_lrbms.py:104-119)
# ----------------------------------------------------------------------------------------------------------

def make_local_bases(data: BlockSwipdgData, basis_size, seed=0, noise=0.05):
    """Return ``[V_i]`` with ``V_i`` of shape ``(N_i, n_i)`` (pyMOR ``(len, dim)`` layout), orthonormal in the
    local energy product.  ``basis_size`` is an int or a per-subdomain sequence (ragged sizes are allowed).

    The span is: the DG shape functions 1, x, y, xy (reference ``discretize...:187-200``), then smooth local
    cosine modes with a small seeded perturbation, Gram-Schmidt'ed (twice) in ``local_energy_dg_product_i``.
    """
    rng = np.random.default_rng(seed)
    S = data.num_subdomains
    sizes = [int(basis_size)] * S if np.isscalar(basis_size) else [int(b) for b in basis_size]
    bases = []
    for i in range(S):
        N = sizes[i]
        xy = data.dof_coords[i]
        lo, hi = xy.min(axis=0), xy.max(axis=0)
        xi = (xy - lo) / (hi - lo)
        cand = [data.shape_functions[i][k] for k in range(min(4, N))]
        modes = sorted(((k, l) for k in range(0, 12) for l in range(0, 12) if k + l > 0 and not (k, l) in ((1, 0), (0, 1))),
                       key=lambda kl: (kl[0] + kl[1], kl))
        for (k, l) in modes:
            if len(cand) >= N:
                break
            v = np.cos(np.pi * k * xi[:, 0]) * np.cos(np.pi * l * xi[:, 1])
            v = v + noise * rng.standard_normal(v.shape) * np.abs(v).max()
            cand.append(v)
        V = np.array(cand[:N]).reshape(N, int(data.n[i]))
        E = data.energy[i]
        for _ in range(2):
            for a in range(N):
                for b in range(a):
                    V[a] -= (V[a] @ (E @ V[b])) * V[b]
                V[a] /= np.sqrt(V[a] @ (E @ V[a]))
        bases.append(np.ascontiguousarray(V))
    return bases
