"""GPU VectorArray backing.

Replaces ``DuneXTVectorSpace`` / ``ListVectorArray`` of ``IstlDenseVectorDouble`` -- one Python object and one C++
``dot`` / ``axpy`` / ``scal`` call per vector (reference ``discretize_elliptic_block_swipdg.py:11,51,85-88,118-120,
174,185,190,223``; ``estimators.py:15,19-23``) -- by one HBM buffer per array.

Layout: **dof-major**.  Element (dof ``d``, vector ``a``) lives at ``buf[d, a]`` of a ``(dim, capacity)`` FP64
buffer (row stride ``ld``), so the SpMM / projection kernels read all vectors of a dof with one coalesced access.
The pyMOR convention ``(len, dim)`` is the transpose and only appears at the ``to_numpy()`` / ``from_data``
boundary.  All arithmetic goes through ``liblrbms_sm100`` (C ABI, ``include/lrbms_sm100.h``); torch only owns the
memory.  There is no CPU fallback.
"""
from __future__ import annotations

import numbers

import numpy as np

from ._lib import Handle, current_stream_ptr, ptr, host_f64, host_i32


def _torch():
    import torch
    return torch


def _round_cap(n):
    return max(4, (int(n) + 3) // 4 * 4)


class GpuVectorSpace:
    """``DuneXTVectorSpace`` / ``NumpyVectorSpace`` stand-in: dimension + id."""

    def __init__(self, dim, id_=None):
        self.dim, self.id = int(dim), id_

    # -- factories (reference discretize...:85,87,150,185,190)
    def empty(self, reserve=0):
        return GpuVectorArray(self, None, 0, reserve=reserve)

    def zeros(self, count=1):
        torch = _torch()
        buf = torch.zeros((self.dim, _round_cap(count)), dtype=torch.float64, device='cuda')
        return GpuVectorArray(self, buf, count)

    def make_array(self, data):
        """``data``: ``(len, dim)`` host array (pyMOR layout), or a :class:`GpuVectorArray` (returned as is)."""
        if isinstance(data, GpuVectorArray):
            assert data.space == self
            return data
        return self.from_data(data)

    def from_data(self, data):
        torch = _torch()
        data = np.ascontiguousarray(np.atleast_2d(np.asarray(data, dtype=np.float64)))
        if data.shape[1] != self.dim:
            raise ValueError('data of shape {} does not fit a space of dimension {}'.format(data.shape, self.dim))
        n = data.shape[0]
        buf = torch.zeros((self.dim, _round_cap(n)), dtype=torch.float64, device='cuda')
        if n and self.dim:
            h = Handle.get()
            src = torch.from_numpy(data).cuda()
            h.check(h.lib.lrbms_va_transpose_in(h.h, self.dim, n, ptr(src), ptr(buf), buf.stride(0), current_stream_ptr()))
        return GpuVectorArray(self, buf, n)

    def from_dofmajor(self, tensor, length=None):
        """Wrap a ``(dim, >= length)`` CUDA tensor (may be a column-slice view of a larger buffer) without copying."""
        assert tensor.dim() == 2 and tensor.shape[0] == self.dim and tensor.stride(1) == 1
        arr = GpuVectorArray(self, tensor, tensor.shape[1] if length is None else length)
        arr._is_view = True
        return arr

    def __eq__(self, other):
        return type(other) is GpuVectorSpace and self.dim == other.dim and self.id == other.id

    def __hash__(self):
        return hash((self.dim, self.id))

    def __repr__(self):
        return 'GpuVectorSpace({}, {!r})'.format(self.dim, self.id)


class GpuVectorArray:
    """pyMOR VectorArray interface on a dof-major HBM buffer."""

    def __init__(self, space, buf, length, reserve=0):
        torch = _torch()
        self.space = space
        if buf is None:
            buf = torch.zeros((space.dim, _round_cap(max(reserve, length))), dtype=torch.float64, device='cuda')
        self._buf, self._len = buf, int(length)
        self._is_view = False      # True for column-slice views of a larger slab: those never grow in place

    # -- basic protocol
    def __len__(self):
        return self._len

    @property
    def dim(self):
        return self.space.dim

    @property
    def ld(self):
        return int(self._buf.stride(0))

    @property
    def device_ptr(self):
        return self._buf.data_ptr()

    def dofmajor(self):
        """``(dim, len)`` view of the device buffer."""
        return self._buf[:, :self._len]

    def to_numpy(self):
        """``(len, dim)`` host array (pyMOR layout)."""
        torch = _torch()
        out = torch.empty((self._len, self.dim), dtype=torch.float64, device='cuda')
        if self._len and self.dim:
            h = Handle.get()
            h.check(h.lib.lrbms_va_transpose_out(h.h, self.dim, self._len, ptr(self._buf), self.ld, ptr(out),
                                                 current_stream_ptr()))
        return out.cpu().numpy()

    data = property(to_numpy)

    def copy(self):
        torch = _torch()
        buf = torch.zeros((self.dim, _round_cap(self._len)), dtype=torch.float64, device='cuda')
        new = GpuVectorArray(self.space, buf, self._len)
        new._copy_cols_from(self, None, 0)
        return new

    def _copy_cols_from(self, other, src, dst0):
        n = len(other) if src is None else len(src)
        if n == 0 or self.dim == 0:
            return
        h = Handle.get()
        for c0 in range(0, n, 256):
            cnt = min(256, n - c0)
            if src is None:
                # identity mapping of a chunk: shift both pointers
                h.check(h.lib.lrbms_va_copy_cols(h.h, self.dim, cnt, None, other.device_ptr + 8 * c0, other.ld,
                                                 ptr(self._buf), self.ld, dst0 + c0, current_stream_ptr()))
            else:
                idx = host_i32(src[c0:c0 + cnt])
                h.check(h.lib.lrbms_va_copy_cols(h.h, self.dim, cnt, ptr(idx), other.device_ptr, other.ld,
                                                 ptr(self._buf), self.ld, dst0 + c0, current_stream_ptr()))

    def _reserve(self, n):
        torch = _torch()
        if n <= self._buf.shape[1] and not self._is_view:
            return
        cap = _round_cap(max(n, 2 * self._len))
        buf = torch.zeros((self.dim, cap), dtype=torch.float64, device='cuda')
        old = GpuVectorArray(self.space, self._buf, self._len)
        self._buf, self._is_view = buf, False
        self._copy_cols_from(old, None, 0)

    def append(self, other, remove_from_other=False):
        assert other.space == self.space, 'append: space mismatch'
        n0 = self._len
        self._reserve(n0 + len(other))
        self._len = n0 + len(other)
        self._copy_cols_from(other, None, n0)

    def __getitem__(self, ind):
        idx = np.arange(self._len)[ind]
        idx = np.atleast_1d(idx)
        new = GpuVectorArray(self.space, None, len(idx))
        new._copy_cols_from(self, idx, 0)
        return new

    def __delitem__(self, ind):
        keep = np.delete(np.arange(self._len), ind)
        tmp = self[keep]
        self._buf, self._len = tmp._buf, tmp._len

    # -- arithmetic
    def _alpha(self, alpha):
        a = np.atleast_1d(np.asarray(alpha, dtype=np.float64)).ravel()
        if a.size not in (1, self._len):
            raise ValueError('alpha must be a scalar or have one entry per vector')
        return host_f64(a)

    def scal(self, alpha):
        if self._len == 0:
            return
        h = Handle.get()
        a = self._alpha(alpha)
        for c0 in range(0, self._len, 256):
            cnt = min(256, self._len - c0)
            ac = a if a.size == 1 else host_f64(a[c0:c0 + cnt])
            h.check(h.lib.lrbms_va_scal(h.h, self.dim, cnt, ptr(ac), ac.size, self.device_ptr + 8 * c0, self.ld,
                                        current_stream_ptr()))

    def axpy(self, alpha, x):
        assert x.space == self.space, 'axpy: space mismatch'
        if len(x) not in (1, self._len):
            raise ValueError('axpy: len(x) must be 1 or len(self)')
        if self._len == 0:
            return
        h = Handle.get()
        a = self._alpha(alpha)
        for c0 in range(0, self._len, 256):
            cnt = min(256, self._len - c0)
            ac = a if a.size == 1 else host_f64(a[c0:c0 + cnt])
            xp = x.device_ptr + (8 * c0 if len(x) > 1 else 0)
            h.check(h.lib.lrbms_va_axpy(h.h, self.dim, cnt, ptr(ac), ac.size, xp, x.ld, 1 if len(x) == 1 else cnt,
                                        self.device_ptr + 8 * c0, self.ld, current_stream_ptr()))

    def dot(self, other):
        """``(len(self), len(other))`` Gram block -- one fused projection launch with the identity operator."""
        from .kernels import project_once
        assert other.space == self.space, 'dot: space mismatch'
        return project_once(None, self, other)

    def pairwise_dot(self, other):
        torch = _torch()
        assert other.space == self.space and len(other) == self._len
        out = torch.zeros(self._len, dtype=torch.float64, device='cuda')
        h = Handle.get()
        for c0 in range(0, self._len, 256):
            cnt = min(256, self._len - c0)
            h.check(h.lib.lrbms_va_pairwise_dot(h.h, self.dim, cnt, self.device_ptr + 8 * c0, self.ld,
                                                other.device_ptr + 8 * c0, other.ld, out.data_ptr() + 8 * c0,
                                                current_stream_ptr()))
        return out.cpu().numpy()

    def lincomb(self, coefficients):
        """``coefficients``: ``(n_out, len(self))`` -> array of ``n_out`` vectors (pyMOR semantics)."""
        torch = _torch()
        if isinstance(coefficients, np.ndarray) or not hasattr(coefficients, 'data_ptr'):
            c = np.atleast_2d(np.asarray(coefficients, dtype=np.float64))
            ct = torch.from_numpy(np.ascontiguousarray(c.T)).cuda()          # (len, n_out) row-major
        else:
            ct = coefficients.t().contiguous() if coefficients.dim() == 2 else coefficients.reshape(-1, 1).contiguous()
        assert ct.shape[0] == self._len, 'lincomb: need one coefficient per vector'
        n_out = ct.shape[1]
        new = GpuVectorArray(self.space, None, n_out)
        if n_out == 0 or self.dim == 0:
            return new
        if self._len == 0:
            return new
        h = Handle.get()
        chunk = max(1, min(n_out, (48 * 1024 // 8) // max(1, self._len)))
        for j0 in range(0, n_out, chunk):
            cnt = min(chunk, n_out - j0)
            h.check(h.lib.lrbms_va_lincomb(h.h, self.dim, self._len, cnt, self.device_ptr, self.ld,
                                           ct.data_ptr() + 8 * j0, ct.stride(0), new.device_ptr + 8 * j0, new.ld,
                                           current_stream_ptr()))
        return new

    def l2_norm(self):
        return np.sqrt(np.maximum(self.pairwise_dot(self), 0.0))

    def l2_norm2(self):
        return self.pairwise_dot(self)

    def __sub__(self, other):
        new = self.copy()
        new.axpy(-1.0, other)
        return new

    def __add__(self, other):
        new = self.copy()
        new.axpy(1.0, other)
        return new

    def __mul__(self, alpha):
        assert isinstance(alpha, numbers.Number)
        new = self.copy()
        new.scal(alpha)
        return new

    def is_zero(self):
        return not bool((self.dofmajor() != 0).any().item())


class BlockVectorSpace:
    """``BlockVectorSpace`` (reference ``reductor.py:40,42``: ``.subspaces[i].id``)."""

    def __init__(self, subspaces, id_=None):
        self.subspaces = list(subspaces)
        self.id = id_
        self.dim = sum(s.dim for s in self.subspaces)

    def zeros(self, count=1):
        return BlockVectorArray([s.zeros(count) for s in self.subspaces], self)

    def empty(self, reserve=0):
        return BlockVectorArray([s.empty(reserve) for s in self.subspaces], self)

    def make_array(self, blocks):
        blocks = list(blocks)
        assert len(blocks) == len(self.subspaces)
        return BlockVectorArray([s.make_array(b) for s, b in zip(self.subspaces, blocks)], self)

    def from_data(self, data):
        data = np.atleast_2d(np.asarray(data, dtype=np.float64))
        offs = np.cumsum([0] + [s.dim for s in self.subspaces])
        return BlockVectorArray([s.from_data(data[:, offs[k]:offs[k + 1]]) for k, s in enumerate(self.subspaces)], self)

    def __eq__(self, other):
        return isinstance(other, BlockVectorSpace) and self.id == other.id and self.subspaces == other.subspaces

    def __hash__(self):
        return hash((self.id, self.dim))

    def __repr__(self):
        return 'BlockVectorSpace({} subspaces, dim {}, {!r})'.format(len(self.subspaces), self.dim, self.id)


class BlockVectorArray:
    """``BlockVectorArray`` with the ``_blocks`` accessor the reference relies on (``discretize...:88,118``)."""

    def __init__(self, blocks, space):
        self._blocks = list(blocks)
        self.space = space
        assert len(self._blocks) == len(space.subspaces)
        assert len({len(b) for b in self._blocks}) <= 1, 'all blocks need the same length'

    def __len__(self):
        return len(self._blocks[0]) if self._blocks else 0

    @property
    def dim(self):
        return self.space.dim

    def block(self, k):
        return self._blocks[k]

    def to_numpy(self):
        return np.hstack([b.to_numpy() for b in self._blocks]) if self._blocks else np.zeros((0, 0))

    data = property(to_numpy)

    def copy(self):
        return BlockVectorArray([b.copy() for b in self._blocks], self.space)

    def append(self, other, remove_from_other=False):
        for b, o in zip(self._blocks, other._blocks):
            b.append(o)

    def scal(self, alpha):
        for b in self._blocks:
            b.scal(alpha)

    def axpy(self, alpha, x):
        for b, o in zip(self._blocks, x._blocks):
            b.axpy(alpha, o)

    def dot(self, other):
        return sum(b.dot(o) for b, o in zip(self._blocks, other._blocks))

    def pairwise_dot(self, other):
        return sum(b.pairwise_dot(o) for b, o in zip(self._blocks, other._blocks))

    def lincomb(self, coefficients):
        return BlockVectorArray([b.lincomb(coefficients) for b in self._blocks], self.space)

    def l2_norm(self):
        return np.sqrt(sum(b.l2_norm2() for b in self._blocks))

    def __sub__(self, other):
        return BlockVectorArray([b - o for b, o in zip(self._blocks, other._blocks)], self.space)

    def __add__(self, other):
        return BlockVectorArray([b + o for b, o in zip(self._blocks, other._blocks)], self.space)

    def __getitem__(self, ind):
        return BlockVectorArray([b[ind] for b in self._blocks], self.space)

    def is_zero(self):
        return all(b.is_zero() for b in self._blocks)


class ReducedVectorArray:
    """Solutions of the reduced model: ``(len, n_red)`` row-major on the device, one row per parameter.

    This is the pyMOR ``(len, dim)`` layout itself (``NumpyVectorArray`` in the reference) because the online
    kernels write one reduced solution per parameter contiguously."""

    def __init__(self, tensor, block_dims=None):
        assert tensor.dim() == 2
        self._t = tensor
        self.block_dims = list(block_dims) if block_dims is not None else None

    def __len__(self):
        return self._t.shape[0]

    @property
    def dim(self):
        return self._t.shape[1]

    @property
    def device_tensor(self):
        return self._t

    def to_numpy(self):
        return self._t.cpu().numpy()

    data = property(to_numpy)

    def __getitem__(self, ind):
        t = self._t[ind]
        return ReducedVectorArray(t.reshape(-1, self._t.shape[1]), self.block_dims)

    def copy(self):
        return ReducedVectorArray(self._t.clone(), self.block_dims)
