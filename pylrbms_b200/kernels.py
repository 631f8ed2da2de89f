"""Thin host helpers over the C ABI: device CSR storage, descriptor builders, one-shot SpMM / projection calls.

Everything numerical happens inside ``liblrbms_sm100`` (``include/lrbms_sm100.h``); this module only packs
pointers.  One-shot helpers create a plan, run it and drop it -- the reductor batches all of its blocks into a
single plan instead (:mod:`pylrbms_b200.reductor`).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from ._lib import Handle, ProjectDesc, SpmmDesc, make_project_plan, make_spmm_plan


def _torch():
    import torch
    return torch


class DeviceCsr:
    """CSR matrix resident in HBM: ``rowptr`` / ``colind`` int32, ``values`` float64.

    Replaces ``IstlRowMajorSparseMatrixDouble`` (reference ``discretize_elliptic_block_swipdg.py:10-14``)."""

    def __init__(self, matrix):
        torch = _torch()
        M = sp.csr_matrix(matrix)
        if not M.has_sorted_indices:
            M = M.copy()
            M.sort_indices()
        self.shape = tuple(int(s) for s in M.shape)
        self.nnz = int(M.nnz)
        self.rowptr = torch.from_numpy(np.ascontiguousarray(M.indptr, dtype=np.int32)).cuda()
        self.colind = torch.from_numpy(np.ascontiguousarray(M.indices if M.nnz else np.zeros(1), dtype=np.int32)).cuda()
        self.values = torch.from_numpy(np.ascontiguousarray(M.data if M.nnz else np.zeros(1), dtype=np.float64)).cuda()
        self._host = M
        self._T = None
        self._symmetric = None

    @classmethod
    def from_device(cls, rowptr, colind, values, shape):
        """Wrap CSR arrays that already live in HBM (int32 ``rowptr`` / ``colind`` with sorted column indices, float64
        ``values``); the host copy is made only if somebody asks for it."""
        self = cls.__new__(cls)
        self.shape = tuple(int(s) for s in shape)
        self.nnz = int(values.numel())
        self.rowptr, self.colind, self.values = rowptr, colind, values
        self._host = None
        self._T = None
        self._symmetric = None
        return self

    @property
    def host(self):
        """The host CSR this was uploaded from (kept for the transposed upload and for inspection)."""
        if self._host is None:
            self._host = sp.csr_matrix((self.values.cpu().numpy(), self.colind.cpu().numpy(), self.rowptr.cpu().numpy()),
                                       shape=self.shape)
        return self._host

    @property
    def T(self):
        if self._T is None:
            self._T = DeviceCsr(self._host.T.tocsr())
            self._T._T = self
        return self._T

    @property
    def symmetric(self):
        """True if the matrix equals its transpose up to rounding of the assembly (checked once, on the host copy)."""
        if self._symmetric is None:
            M = self._host
            if M.shape[0] != M.shape[1]:
                self._symmetric = False
            elif M.nnz == 0:
                self._symmetric = True
            else:
                D = abs(M - M.T)
                self._symmetric = bool(D.max() <= 1e-13 * abs(M).max())
        return self._symmetric

    @property
    def device_bytes(self):
        return 4 * (self.shape[0] + 1) + 12 * self.nnz


def spmm_desc(csr, V_ptr, ldv, N, W_ptr, ldw):
    return SpmmDesc(csr.rowptr.data_ptr(), csr.colind.data_ptr(), csr.values.data_ptr(), csr.shape[0], csr.shape[1],
                    V_ptr, ldv, N, W_ptr, ldw)


def project_desc(csr, n_rows, VL_ptr, ldl, NL, VR_ptr, ldr, NR, out_ptr, ldo, alpha=1.0, symmetric=False):
    """``csr=None`` selects the identity operator (Gram matrix ``VL^T VR``).  ``symmetric=True`` asserts that the
    operator is symmetric and ``VL`` / ``VR`` are the same array (only the lower output chunks are computed)."""
    sym = 1 if (symmetric and VL_ptr == VR_ptr and ldl == ldr and NL == NR) else 0
    if csr is None:
        return ProjectDesc(None, None, None, n_rows, n_rows, VL_ptr, ldl, NL, VR_ptr, ldr, NR, out_ptr, ldo, float(alpha), sym, 0)
    assert csr.shape[0] == n_rows
    return ProjectDesc(csr.rowptr.data_ptr(), csr.colind.data_ptr(), csr.values.data_ptr(), csr.shape[0], csr.shape[1],
                       VL_ptr, ldl, NL, VR_ptr, ldr, NR, out_ptr, ldo, float(alpha), sym, 0)


def spmm_once(csr, V, range_space):
    """``W = A V`` for a :class:`GpuVectorArray` ``V``; returns a new array in ``range_space``."""
    from .vectorarray import GpuVectorArray
    W = GpuVectorArray(range_space, None, len(V))
    if len(V) == 0 or csr.shape[0] == 0:
        return W
    h = Handle.get()
    plan = make_spmm_plan(h, [spmm_desc(csr, V.device_ptr, V.ld, len(V), W.device_ptr, W.ld)], [csr, V, W])
    plan.run()
    plan.destroy()
    return W


def project_once(csr, VL, VR, alpha=1.0):
    """``alpha * VL^T (A VR)`` as a host ``(len(VL), len(VR))`` array (``csr=None``: ``VL^T VR``)."""
    torch = _torch()
    NL, NR = len(VL), len(VR)
    if NL == 0 or NR == 0:
        return np.zeros((NL, NR))
    out = torch.zeros((NL, NR), dtype=torch.float64, device='cuda')
    if VL.dim == 0:
        return out.cpu().numpy()
    h = Handle.get()
    d = project_desc(csr, VL.dim, VL.device_ptr, VL.ld, NL, VR.device_ptr, VR.ld, NR, out.data_ptr(), NR, alpha)
    plan = make_project_plan(h, [d], [csr, VL, VR, out])
    plan.run()
    plan.destroy()
    return out.cpu().numpy()


def pcg_solve(csr, b, x0=None, rtol=1e-12, max_iter=None):
    """Solve ``A x = b`` for a symmetric positive definite :class:`DeviceCsr` ``A`` with the Jacobi-preconditioned CG of
    ``lrbms_pcg_solve``.  ``b`` / ``x0``: 1-D float64 CUDA tensors.  Returns ``(x, iterations, relative residual)``."""
    import ctypes as C
    torch = _torch()
    n = csr.shape[0]
    assert csr.shape[0] == csr.shape[1] == b.numel()
    h = Handle.get()
    x = torch.zeros(n, dtype=torch.float64, device='cuda') if x0 is None else x0.clone().contiguous()
    nbytes = C.c_size_t()
    h.check(h.lib.lrbms_pcg_workspace_bytes(h.h, n, C.byref(nbytes)))
    ws = torch.empty(max(1, nbytes.value // 8), dtype=torch.float64, device='cuda')
    iters, relres = C.c_int32(), C.c_double()
    from ._lib import current_stream_ptr
    h.check(h.lib.lrbms_pcg_solve(h.h, n, csr.rowptr.data_ptr(), csr.colind.data_ptr(), csr.values.data_ptr(),
                                  b.contiguous().data_ptr(), x.data_ptr(), float(rtol),
                                  int(max_iter if max_iter is not None else max(100, 10 * n)), C.byref(iters), C.byref(relres),
                                  ws.data_ptr(), nbytes.value, current_stream_ptr()))
    return x, iters.value, relres.value
