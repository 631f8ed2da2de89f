"""TEST INFRASTRUCTURE ONLY -- CPU oracle, part 1: the VectorArray / Operator algebra the reference runs on.

**Parity status**: the reference's own tests hold no golden vector for this path and the pyMOR
fork + DUNE it executes on cannot be installed here.  This file restates, in plain NumPy/SciPy, the *published* pyMOR-0.5
semantics the reference relies on and is anchored on the reference's call sites, which
are cited per function -- **this layer stays unpinned** (a third-party dependency absent from ``/root/reference``).  The
reference's OWN files for the path are pinned: ``oracle/reference_run.py`` executes ``estimators.py`` / ``reductor.py`` /
``online_enrichment.py`` unmodified on top of this layer and their outputs are committed as
``tests/golden/reference_run__*.npz``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this file; the product package never does.

Array convention (Appendix A.1): a VectorArray of length L in a space of dimension n is an ``(L, n)`` array.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# ----------------------------------------------------------------------------------------------------------
#  vector spaces / arrays
# ----------------------------------------------------------------------------------------------------------


class Space:
    def __init__(self, dim, id_=None):
        self.dim, self.id = int(dim), id_

    def zeros(self, count=1):
        return VA(np.zeros((count, self.dim)), self)

    def empty(self, reserve=0):
        return VA(np.zeros((0, self.dim)), self)

    def make_array(self, data):
        return VA(np.atleast_2d(np.asarray(data, dtype=float)), self)

    from_data = make_array

    def __eq__(self, other):
        return isinstance(other, Space) and not isinstance(other, BlockSpace) and self.dim == other.dim and self.id == other.id

    def __hash__(self):
        return hash((self.dim, self.id))


class VA:
    """NumpyVectorArray / ListVectorArray stand-in: ``data`` is ``(len, dim)``."""

    def __init__(self, data, space):
        self.data = np.asarray(data, dtype=float)
        assert self.data.ndim == 2 and self.data.shape[1] == space.dim
        self.space = space

    def __len__(self):
        return self.data.shape[0]

    @property
    def dim(self):
        return self.space.dim

    def copy(self):
        return VA(self.data.copy(), self.space)

    def append(self, other):
        self.data = np.vstack([self.data, other.data])

    def scal(self, alpha):
        self.data = self.data * alpha

    def axpy(self, alpha, x):
        self.data = self.data + alpha * x.data

    def dot(self, other):
        return self.data @ other.data.T

    def pairwise_dot(self, other):
        return np.einsum('ij,ij->i', self.data, other.data)

    def lincomb(self, coefficients):
        return VA(np.atleast_2d(coefficients) @ self.data, self.space)

    def l2_norm(self):
        return np.linalg.norm(self.data, axis=1)

    def __sub__(self, other):
        return VA(self.data - other.data, self.space)

    def __getitem__(self, ind):
        d = self.data[ind]
        return VA(np.atleast_2d(d), self.space)

    def is_zero(self):
        return not np.any(self.data)


class _ZeroVA(VA):
    """Shared, immutable all-zero array (allocation saver of the oracle; any in-place use is a bug and raises)."""

    def __init__(self, space, count):
        z = np.zeros((count, space.dim))
        z.setflags(write=False)
        super().__init__(z, space)

    def _immutable(self, *a, **k):
        raise RuntimeError('shared zero array must not be modified in place')

    append = scal = axpy = _immutable

    def copy(self):
        return VA(np.zeros(self.data.shape), self.space)

    def is_zero(self):
        return True


_ZERO_CACHE = {}


def shared_zeros(space, count):
    """An all-zero array of ``space`` that may be shared between callers (read-only)."""
    if isinstance(space, BlockSpace):
        return BlockVA([shared_zeros(s, count) for s in space.subspaces], space)
    key = (space.dim, space.id, count)
    z = _ZERO_CACHE.get(key)
    if z is None:
        if len(_ZERO_CACHE) > 8192:
            _ZERO_CACHE.clear()
        z = _ZERO_CACHE[key] = _ZeroVA(space, count)
    return z


class BlockSpace(Space):
    def __init__(self, subspaces, id_=None):
        self.subspaces = list(subspaces)
        self.id = id_
        self.dim = sum(s.dim for s in self.subspaces)

    def zeros(self, count=1):
        return BlockVA([s.zeros(count) for s in self.subspaces], self)

    def empty(self, reserve=0):
        return BlockVA([s.empty() for s in self.subspaces], self)

    def make_array(self, blocks):
        return BlockVA(list(blocks), self)

    def from_data(self, data):
        data = np.atleast_2d(data)
        offs = np.cumsum([0] + [s.dim for s in self.subspaces])
        return BlockVA([s.from_data(data[:, offs[k]:offs[k + 1]]) for k, s in enumerate(self.subspaces)], self)

    def __eq__(self, other):
        return isinstance(other, BlockSpace) and self.id == other.id and len(self.subspaces) == len(other.subspaces) \
            and all(a == b for a, b in zip(self.subspaces, other.subspaces))

    def __hash__(self):
        return hash((self.id, self.dim))


class BlockVA:
    def __init__(self, blocks, space):
        self._blocks = list(blocks)
        self.space = space

    def __len__(self):
        return len(self._blocks[0])

    @property
    def dim(self):
        return self.space.dim

    @property
    def data(self):
        return np.hstack([b.data for b in self._blocks])

    def copy(self):
        return BlockVA([b.copy() for b in self._blocks], self.space)

    def append(self, other):
        for b, o in zip(self._blocks, other._blocks):
            b.append(o)

    def scal(self, alpha):
        self._blocks = [b if isinstance(b, _ZeroVA) else _scaled(b, alpha) for b in self._blocks]

    def axpy(self, alpha, x):
        self._blocks = [b if isinstance(o, _ZeroVA) else _axpyed(b, alpha, o) for b, o in zip(self._blocks, x._blocks)]

    def dot(self, other):
        return sum(b.dot(o) for b, o in zip(self._blocks, other._blocks))

    def pairwise_dot(self, other):
        return sum(b.pairwise_dot(o) for b, o in zip(self._blocks, other._blocks))

    def lincomb(self, coefficients):
        return BlockVA([b.lincomb(coefficients) for b in self._blocks], self.space)

    def l2_norm(self):
        return np.sqrt(sum(b.l2_norm() ** 2 for b in self._blocks))

    def __sub__(self, other):
        return BlockVA([b - o for b, o in zip(self._blocks, other._blocks)], self.space)

    def __getitem__(self, ind):
        return BlockVA([b[ind] for b in self._blocks], self.space)

    def is_zero(self):
        return all(b.is_zero() for b in self._blocks)


def _scaled(b, alpha):
    if isinstance(b, BlockVA):
        new = BlockVA(list(b._blocks), b.space)
        new.scal(alpha)
        return new
    return VA(b.data * alpha, b.space)


def _axpyed(b, alpha, o):
    if isinstance(b, BlockVA):
        new = BlockVA(list(b._blocks), b.space)
        new.axpy(alpha, o)
        return new
    return VA(b.data + alpha * o.data, b.space)


NUMBER_SPACE = Space(1, 'SCALARS')

# ----------------------------------------------------------------------------------------------------------
#  parameter functionals (Appendix A.8)
# ----------------------------------------------------------------------------------------------------------


class ExpressionParameterFunctional:
    """``ExpressionParameterFunctional('diffusion', parameter_type)`` (reference OS2015_academic_problem.py:43-44)."""
    _ns = {'sin': np.sin, 'cos': np.cos, 'exp': np.exp, 'sqrt': np.sqrt, 'pi': np.pi, 'abs': np.abs,
           'min': np.minimum, 'max': np.maximum, 'log': np.log, 'tan': np.tan}

    def __init__(self, expression, parameter_type=None):
        self.expression, self.parameter_type = expression, parameter_type

    def evaluate(self, mu=None):
        env = dict(self._ns)
        for k, v in (mu or {}).items():
            v = np.asarray(v, dtype=float).ravel()
            env[k] = float(v[0]) if v.size == 1 else v
        return float(eval(self.expression, {'__builtins__': {}}, env))     # noqa: S307


class ProductParameterFunctional:
    """Product of evaluations (reference discretize_elliptic_block_swipdg.py:757-759)."""

    def __init__(self, factors):
        self.factors = list(factors)

    def evaluate(self, mu=None):
        out = 1.0
        for f in self.factors:
            out = out * (f.evaluate(mu) if hasattr(f, 'evaluate') else f)
        return out


def _coeff(c, mu):
    return c.evaluate(mu) if hasattr(c, 'evaluate') else c


# ----------------------------------------------------------------------------------------------------------
#  operators
# ----------------------------------------------------------------------------------------------------------


class Operator:
    linear = True
    name = None

    def apply2(self, V, U, mu=None):
        """Appendix A.2: ``V.dot(op.apply(U, mu))`` -> ``(len(V), len(U))``."""
        return V.dot(self.apply(U, mu=mu))

    def pairwise_apply2(self, V, U, mu=None):
        """Appendix A.2: ``V.pairwise_dot(op.apply(U, mu))`` (used at reference estimators.py:71-85)."""
        return V.pairwise_dot(self.apply(U, mu=mu))

    def assemble(self, mu=None):
        return self

    def with_(self, **kw):
        import copy
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new


class MatrixOperator(Operator):
    """DuneXTMatrixOperator / NumpyMatrixOperator stand-in: a sparse or dense matrix, range x source.

    ``per_vector=True`` walks the array one vector at a time -- one ``mv`` per basis vector, as the reference's
    ListVectorArray does -- and is what the CPU baseline times."""
    per_vector = False

    def __init__(self, matrix, source_id=None, range_id=None, name=None):
        self.matrix = matrix
        self.sparse = sp.issparse(matrix)
        self.source = Space(matrix.shape[1], source_id)
        self.range = Space(matrix.shape[0], range_id)
        self.name = name

    def apply(self, U, mu=None):
        if MatrixOperator.per_vector:
            out = np.empty((len(U), self.range.dim))
            for k in range(len(U)):
                out[k] = self.matrix @ U.data[k]
            return VA(out, self.range)
        return VA((self.matrix @ U.data.T).T, self.range)

    def apply_transpose(self, V, mu=None):
        return VA((self.matrix.T @ V.data.T).T, self.source)

    def apply_inverse(self, V, mu=None):
        """Appendix A.7: dense -> ``numpy.linalg.solve``."""
        M = self.matrix.toarray() if self.sparse else self.matrix
        return VA(np.linalg.solve(M, V.data.T).T, self.source)

    @property
    def T(self):
        return MatrixOperator(self.matrix.T.tocsr() if self.sparse else self.matrix.T,
                              source_id=self.range.id, range_id=self.source.id)

    def as_source_array(self, mu=None):
        M = self.matrix.toarray() if self.sparse else self.matrix
        return VA(M, self.source)


class VectorFunctional(Operator):
    """``VectorFunctional(array)``: ``U -> U . v`` (reference discretize_elliptic_block_swipdg.py:522-526,742)."""

    def __init__(self, array):
        self._array = array
        self.source = array.space
        self.range = NUMBER_SPACE

    def apply(self, U, mu=None):
        return VA(U.dot(self._array), NUMBER_SPACE)

    def as_source_array(self, mu=None):
        return self._array.copy()

    as_vector = as_source_array


class VectorArrayOperator(Operator):
    """Appendix A.3: what a projected functional becomes (``transposed=True``: source = array.space)."""

    def __init__(self, array, transposed=False):
        self._array, self.transposed = array, transposed
        if transposed:
            self.source, self.range = array.space, Space(len(array))
        else:
            self.source, self.range = Space(len(array)), array.space

    def apply(self, U, mu=None):
        if self.transposed:
            return VA(U.dot(self._array), self.range)
        return self._array.lincomb(U.data)

    def as_source_array(self, mu=None):
        assert self.transposed
        return self._array.copy()


class LincombOperator(Operator):
    def __init__(self, operators, coefficients, name=None, solver_options=None):
        self.operators, self.coefficients, self.name = list(operators), list(coefficients), name
        self.source, self.range = self.operators[0].source, self.operators[0].range

    def evaluate_coefficients(self, mu):
        return [_coeff(c, mu) for c in self.coefficients]

    def apply(self, U, mu=None):
        cs = self.evaluate_coefficients(mu)
        R = self.operators[0].apply(U, mu=mu)
        R.scal(cs[0])
        for op, c in zip(self.operators[1:], cs[1:]):
            R.axpy(c, op.apply(U, mu=mu))
        return R

    def assemble(self, mu=None):
        """Appendix A.6: ``M = c_0 M_0; M += c_q M_q`` left to right."""
        cs = self.evaluate_coefficients(mu)
        ops = [op.assemble(mu) for op in self.operators]
        if all(isinstance(o, MatrixOperator) for o in ops):
            M = ops[0].matrix * cs[0]
            for o, c in zip(ops[1:], cs[1:]):
                M = M + o.matrix * c
            return MatrixOperator(M, source_id=self.source.id, range_id=self.range.id, name=self.name)
        if all(isinstance(o, VectorArrayOperator) for o in ops):
            A = ops[0]._array.copy()
            A.scal(cs[0])
            for o, c in zip(ops[1:], cs[1:]):
                A.axpy(c, o._array)
            return VectorArrayOperator(A, transposed=ops[0].transposed)
        return self

    def as_source_array(self, mu=None):
        cs = self.evaluate_coefficients(mu)
        R = self.operators[0].as_source_array(mu)
        R.scal(cs[0])
        for op, c in zip(self.operators[1:], cs[1:]):
            R.axpy(c, op.as_source_array(mu))
        return R

    as_vector = as_source_array

    def apply_inverse(self, V, mu=None):
        return self.assemble(mu).apply_inverse(V, mu=mu)


class Concatenation(Operator):
    """List form ``Concatenation([A, B, C])`` = A o B o C (fork-only, reference discretize...:356,734,745,748)."""

    def __init__(self, operators, name=None):
        self.operators, self.name = list(operators), name
        self.source, self.range = self.operators[-1].source, self.operators[0].range

    def apply(self, U, mu=None):
        for op in reversed(self.operators):
            U = op.apply(U, mu=mu)
        return U

    @property
    def T(self):
        return Concatenation([op.T for op in reversed(self.operators)])


class BlockOperator(Operator):
    def __init__(self, blocks, range_spaces=None, source_spaces=None, name=None, range_id=None, source_id=None):
        blocks = np.asarray(blocks, dtype=object)
        if blocks.ndim == 1:
            blocks = blocks.reshape(self._shape1d(blocks))
        self._blocks = blocks
        nr, ns = blocks.shape
        if range_spaces is None:
            range_spaces = [next(b.range for b in blocks[i, :] if b is not None) for i in range(nr)]
        if source_spaces is None:
            source_spaces = [next(b.source for b in blocks[:, j] if b is not None) for j in range(ns)]
        self.range = BlockSpace(range_spaces, range_id) if self._block_range else range_spaces[0]
        self.source = BlockSpace(source_spaces, source_id) if self._block_source else source_spaces[0]
        self.name = name

    _block_range = True
    _block_source = True

    def _shape1d(self, blocks):
        raise ValueError

    def apply(self, U, mu=None):
        Ub = U._blocks if self._block_source else [U]
        out = []
        nr, ns = self._blocks.shape
        rs = self.range.subspaces if self._block_range else [self.range]
        for i in range(nr):
            acc = None
            for j in range(ns):
                b = self._blocks[i, j]
                if b is not None:
                    W = b.apply(Ub[j], mu=mu)
                    acc = W if acc is None else _axpyed(acc, 1.0, W)
            out.append(acc if acc is not None else shared_zeros(rs[i], len(U)))
        return BlockVA(out, self.range) if self._block_range else out[0]

    @property
    def T(self):
        raise NotImplementedError


class BlockDiagonalOperator(BlockOperator):
    def __init__(self, blocks, name=None, range_id=None, source_id=None):
        n = len(blocks)
        arr = np.full((n, n), None, dtype=object)
        for i, b in enumerate(blocks):
            arr[i, i] = b
        super().__init__(arr, name=name, range_id=range_id, source_id=source_id)


class BlockProjectionOperator(Operator):
    """Picks component ``index`` of a block array (fork-only; reference discretize...:696,704,714)."""

    def __init__(self, block_space, index):
        self.source, self.index = block_space, index
        self.range = block_space.subspaces[index]

    def apply(self, U, mu=None):
        return U._blocks[self.index].copy()

    @property
    def T(self):
        return BlockEmbeddingOperator(self.source, self.index)


class BlockEmbeddingOperator(Operator):
    def __init__(self, block_space, index):
        self.range, self.index = block_space, index
        self.source = block_space.subspaces[index]

    def apply(self, U, mu=None):
        return _embed(self.range, self.index, U.copy())

    @property
    def T(self):
        return BlockProjectionOperator(self.range, self.index)


class BlockRowOperator(BlockOperator):
    """1 x S block operator with a non-block range (fork-only; reference discretize...:705,715)."""
    _block_range = False

    def __init__(self, blocks, source_spaces=None, name=None):
        arr = np.full((1, len(blocks)), None, dtype=object)
        for j, b in enumerate(blocks):
            arr[0, j] = b
        rng = next(b.range for b in blocks if b is not None)
        super().__init__(arr, range_spaces=[rng], source_spaces=source_spaces, name=name)

    @property
    def T(self):
        return BlockColumnOperator([b.T if b is not None else None for b in self._blocks[0, :]],
                                   range_spaces=self.source.subspaces)


class BlockColumnOperator(BlockOperator):
    _block_source = False

    def __init__(self, blocks, range_spaces=None, name=None):
        arr = np.full((len(blocks), 1), None, dtype=object)
        for i, b in enumerate(blocks):
            arr[i, 0] = b
        src = next(b.source for b in blocks if b is not None)
        super().__init__(arr, range_spaces=range_spaces, source_spaces=[src], name=name)

    @property
    def T(self):
        return BlockRowOperator([b.T if b is not None else None for b in self._blocks[:, 0]],
                                source_spaces=self.range.subspaces)


# ----------------------------------------------------------------------------------------------------------
#  projection (Appendix A.3-A.5)
# ----------------------------------------------------------------------------------------------------------


def project(op, range_basis, source_basis):
    """Appendix A.3.  Linear op -> dense ``op.apply2(RB, SB)``; Lincomb keeps its coefficients."""
    if isinstance(op, LincombOperator):
        return LincombOperator([project(o, range_basis, source_basis) for o in op.operators], op.coefficients, name=op.name)
    if range_basis is None:
        if source_basis is None:
            return op
        assert isinstance(op, (VectorFunctional,)) or op.range.dim == 1
        # functional: keep op.apply(SB) as a VectorArrayOperator  (reduced rhs, cf. reference reductor.py:102-118)
        V = op.apply(source_basis)                  # (len(SB), 1)
        return VectorArrayOperator(VA(V.data.T, Space(len(source_basis))), transposed=True)
    M = op.apply2(range_basis, source_basis)
    return MatrixOperator(M, source_id=getattr(op.source, 'id', None), range_id=getattr(op.range, 'id', None), name=op.name)


def _embed(space, index, basis):
    """Zero-embed ``basis`` (an array in ``space.subspaces[index]``) into block ``space``; the zero blocks are shared
    read-only arrays, which keeps the large configurations affordable."""
    return BlockVA([basis if k == index else shared_zeros(sub, len(basis)) for k, sub in enumerate(space.subspaces)], space)


def _touched_source_blocks(op):
    """Source subspaces a chain can depend on: those its right-most selection operator reads.  Every other subspace
    is mapped to exactly zero (and would be dropped by the ``is_zero`` test below), so skipping it changes nothing
    but the run time."""
    while isinstance(op, Concatenation):
        op = op.operators[-1]
    if isinstance(op, BlockRowOperator):
        return [j for j, b in enumerate(op._blocks[0, :]) if b is not None]
    if isinstance(op, BlockProjectionOperator):
        return [op.index]
    return None


def project_system(op, range_bases, source_bases):
    """Appendix A.4: block (i, j) is projected with the bases looked up *by subspace id*.

    For a ``BlockOperator`` the blocks are projected one by one (``None`` stays ``None``).  Any other operator on
    block spaces (the ``Concatenation`` chains of the estimator, reference discretize...:733-770) is projected
    through its definition ``B_i^T op(E_j B_j)`` with zero blocks detected and kept as ``None``.
    """
    if isinstance(op, LincombOperator):
        return LincombOperator([project_system(o, range_bases, source_bases) for o in op.operators],
                               op.coefficients, name=op.name)
    src_block = isinstance(op.source, BlockSpace)
    rng_block = isinstance(op.range, BlockSpace)
    src_sub = op.source.subspaces if src_block else [op.source]
    rng_sub = op.range.subspaces if rng_block else [op.range]
    blocks = np.full((len(rng_sub), len(src_sub)), None, dtype=object)
    if isinstance(op, BlockOperator) and src_block and rng_block:
        for (i, j), b in np.ndenumerate(op._blocks):
            if b is not None:
                blocks[i, j] = project(b, range_bases[rng_sub[i].id], source_bases[src_sub[j].id])
    else:
        touched = _touched_source_blocks(op) if src_block else None
        for j, ss in enumerate(src_sub):
            if touched is not None and j not in touched:
                continue
            SB = source_bases[ss.id]
            W = op.apply(_embed(op.source, j, SB) if src_block else SB)
            if W.is_zero():
                continue
            Wb = W._blocks if rng_block else [W]
            for i, rs in enumerate(rng_sub):
                if rng_block and Wb[i].is_zero():
                    continue
                if rs is NUMBER_SPACE or (rs.id == 'SCALARS'):
                    blocks[i, j] = VectorArrayOperator(VA(Wb[i].data.T, Space(len(SB), ss.id)), transposed=True)
                else:
                    RB = range_bases[rs.id]
                    blocks[i, j] = MatrixOperator(RB.dot(Wb[i]), source_id=ss.id, range_id=rs.id)
    return _ReducedBlockOperator(blocks, [len(range_bases[s.id]) if s.id != 'SCALARS' else 1 for s in rng_sub],
                                 [len(source_bases[s.id]) for s in src_sub], name=op.name)


class _ReducedBlockOperator(Operator):
    """Block operator of dense reduced blocks (``None`` = zero); ``unblock`` turns it into one dense matrix."""

    def __init__(self, blocks, range_dims, source_dims, name=None):
        self._blocks, self.range_dims, self.source_dims, self.name = blocks, list(range_dims), list(source_dims), name
        self.source, self.range = Space(sum(source_dims)), Space(sum(range_dims))

    def dense(self):
        ro = np.cumsum([0] + self.range_dims)
        so = np.cumsum([0] + self.source_dims)
        M = np.zeros((ro[-1], so[-1]))
        for (i, j), b in np.ndenumerate(self._blocks):
            if b is None:
                continue
            M[ro[i]:ro[i + 1], so[j]:so[j + 1]] = b.matrix if isinstance(b, MatrixOperator) else b._array.data
        return M


def unblock(op):
    """Appendix A.5 (fork-only ``pymor.algorithms.system.unblock``, reference reductor.py:7,46,66)."""
    if isinstance(op, LincombOperator):
        return LincombOperator([unblock(o) for o in op.operators], op.coefficients, name=op.name)
    if isinstance(op, _ReducedBlockOperator):
        M = op.dense()
        if op.range.dim == 1 and all(isinstance(b, VectorArrayOperator) for b in op._blocks.ravel() if b is not None):
            return VectorArrayOperator(VA(M, Space(M.shape[1])), transposed=True)
        return MatrixOperator(M, name=op.name)
    if isinstance(op, BlockOperator):
        rs = op.range.subspaces
        ss = op.source.subspaces
        blocks = op._blocks
        return unblock(_ReducedBlockOperator(blocks, [s.dim for s in rs], [s.dim for s in ss], name=op.name))
    return op
