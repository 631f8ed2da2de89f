"""pyMOR-shaped operators on GPU arrays (SURVEY.md section 8b "Python surface to keep").

These mirror the operator classes the reference composes in ``discretize()``
(``discretize_elliptic_block_swipdg.py:50-61`` imports; ``:319-378, 473-507, 596-618, 639-770`` uses) so a
discretization built from them looks like the reference's ``d`` to ``LRBMSReductor`` and ``EllipticEstimator``:
same attributes (``source``, ``range``, ``operators``, ``coefficients``, ``_blocks``, ``.T``), same ``apply`` /
``apply2`` / ``pairwise_apply2`` meaning.  The arithmetic is always a kernel of
``liblrbms_sm100``: SpMM for ``apply``, the fused projection for ``apply2``.

The generic ``apply`` chain is what the reference *executes*; the reductor does not use it for the offline
projection -- it reads the operator *structure* and batches every block into one projection plan
(:mod:`pylrbms_b200.reductor`).
"""
from __future__ import annotations

import copy

import numpy as np

from .kernels import DeviceCsr, project_once, spmm_once
from .parameters import evaluate as _evaluate_coefficient
from .vectorarray import BlockVectorArray, BlockVectorSpace, GpuVectorSpace


class NumpyVectorArray:
    """Tiny host array for scalar-valued results (range of functionals); ``data`` is ``(len, dim)``."""

    def __init__(self, data, space=None):
        self._data = np.atleast_2d(np.asarray(data, dtype=np.float64))
        self.space = space if space is not None else GpuVectorSpace(self._data.shape[1], 'SCALARS')

    def __len__(self):
        return self._data.shape[0]

    @property
    def dim(self):
        return self._data.shape[1]

    @property
    def data(self):
        return self._data

    def to_numpy(self):
        return self._data

    def scal(self, alpha):
        self._data = self._data * np.asarray(alpha, dtype=float).reshape(-1, 1) if np.ndim(alpha) else self._data * alpha

    def axpy(self, alpha, x):
        self._data = self._data + alpha * x._data

    def copy(self):
        return NumpyVectorArray(self._data.copy(), self.space)

    def is_zero(self):
        return not np.any(self._data)


NUMBER_SPACE = GpuVectorSpace(1, 'SCALARS')


class Operator:
    linear = True
    name = None
    source = None
    range = None

    def apply(self, U, mu=None):
        raise NotImplementedError

    def apply2(self, V, U, mu=None):
        """``V.dot(op.apply(U, mu))`` -> ``(len(V), len(U))``."""
        return V.dot(self.apply(U, mu=mu))

    def pairwise_apply2(self, V, U, mu=None):
        """``V.pairwise_dot(op.apply(U, mu))`` -> ``(len(U),)`` (used at reference ``estimators.py:71-85``)."""
        return V.pairwise_dot(self.apply(U, mu=mu))

    def assemble(self, mu=None):
        return self

    def with_(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new


class CsrOperator(Operator):
    """``DuneXTMatrixOperator`` stand-in: one sparse matrix in HBM (reference ``discretize...:333,353,375,473,502,
    670,679,689,725``).  ``apply`` is one SpMM over *all* vectors of ``U`` instead of one ``mv`` per vector."""

    def __init__(self, matrix, source_id=None, range_id=None, name=None, solver_options=None):
        self.csr = matrix if isinstance(matrix, DeviceCsr) else DeviceCsr(matrix)
        self.source = GpuVectorSpace(self.csr.shape[1], source_id)
        self.range = GpuVectorSpace(self.csr.shape[0], range_id)
        self.name = name
        self.solver_options = solver_options
        self.transposed_of = None

    @property
    def matrix(self):
        """Host CSR (inspection only)."""
        return self.csr.host

    def apply(self, U, mu=None):
        assert U.space.dim == self.source.dim
        return spmm_once(self.csr, U, self.range)

    def apply_transpose(self, V, mu=None):
        return spmm_once(self.csr.T, V, self.source)

    def apply2(self, V, U, mu=None):
        return project_once(self.csr, V, U)

    def pairwise_apply2(self, V, U, mu=None):
        return V.pairwise_dot(self.apply(U))

    @property
    def T(self):
        op = CsrOperator(self.csr.T, source_id=self.range.id, range_id=self.source.id,
                         name=None if self.name is None else self.name + '_T')
        op.transposed_of = self
        return op


class VectorFunctional(Operator):
    """``VectorFunctional(array)``: ``U -> U . v`` (reference ``discretize...:522-526,742``)."""

    def __init__(self, array, name=None):
        assert len(array) == 1
        self._array = array
        self.source = array.space
        self.range = NUMBER_SPACE
        self.name = name

    def apply(self, U, mu=None):
        return NumpyVectorArray(U.dot(self._array), NUMBER_SPACE)

    def as_source_array(self, mu=None):
        return self._array.copy()

    as_vector = as_source_array


class LincombOperator(Operator):
    def __init__(self, operators, coefficients, name=None, solver_options=None):
        self.operators, self.coefficients = list(operators), list(coefficients)
        assert len(self.operators) == len(self.coefficients) and self.operators
        self.source, self.range = self.operators[0].source, self.operators[0].range
        self.name, self.solver_options = name, solver_options

    def evaluate_coefficients(self, mu):
        return [_evaluate_coefficient(c, mu) for c in self.coefficients]

    def apply(self, U, mu=None):
        cs = self.evaluate_coefficients(mu)
        R = self.operators[0].apply(U, mu=mu)
        R.scal(cs[0])
        for op, c in zip(self.operators[1:], cs[1:]):
            R.axpy(c, op.apply(U, mu=mu))
        return R

    def as_source_array(self, mu=None):
        cs = self.evaluate_coefficients(mu)
        R = self.operators[0].as_source_array(mu)
        R.scal(cs[0])
        for op, c in zip(self.operators[1:], cs[1:]):
            R.axpy(c, op.as_source_array(mu))
        return R

    as_vector = as_source_array

    def assemble(self, mu=None):
        ops = [op.assemble(mu) for op in self.operators]
        if all(hasattr(o, 'lincomb_assemble') for o in ops):
            return ops[0].lincomb_assemble(ops, self.evaluate_coefficients(mu), name=self.name)
        return self

    def apply_inverse(self, V, mu=None):
        A = self.assemble(mu)
        if A is self:
            raise NotImplementedError('apply_inverse of an unassembled LincombOperator (the fine-scale solve is '
                                      'outside the LRBMS hot path)')
        return A.apply_inverse(V, mu=mu)


class Concatenation(Operator):
    """List form ``Concatenation([A, B, C]) = A o B o C`` (fork-only; reference ``discretize...:356,734,745,748``)."""

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

    def flat(self):
        out = []
        for op in self.operators:
            out.extend(op.flat() if isinstance(op, Concatenation) else [op])
        return out


class BlockOperator(Operator):
    """``BlockOperator`` with the ``_blocks`` object array (reference ``reductor.py:41,58``; built at
    ``discretize...:338,505-506``)."""
    _block_range = True
    _block_source = True

    def __init__(self, blocks, range_spaces=None, source_spaces=None, name=None, range_id=None, source_id=None,
                 dof_communicator=None):
        blocks = np.asarray(blocks, dtype=object)
        assert blocks.ndim == 2
        self._blocks = blocks
        nr, ns = blocks.shape
        if range_spaces is None:
            range_spaces = [next(b.range for b in blocks[i, :] if b is not None) for i in range(nr)]
        if source_spaces is None:
            source_spaces = [next(b.source for b in blocks[:, j] if b is not None) for j in range(ns)]
        self.range = BlockVectorSpace(range_spaces, range_id) if self._block_range else range_spaces[0]
        self.source = BlockVectorSpace(source_spaces, source_id) if self._block_source else source_spaces[0]
        self.name = name
        self._nonzero = [(int(i), int(j)) for i, j in zip(*np.nonzero(blocks != None))]     # noqa: E711

    def nonzero_blocks(self):
        """``[(i, j, block)]`` of the stored blocks (the block array of an S x S system is almost empty)."""
        return [(i, j, self._blocks[i, j]) for (i, j) in self._nonzero]

    @property
    def num_range_blocks(self):
        return self._blocks.shape[0]

    @property
    def num_source_blocks(self):
        return self._blocks.shape[1]

    def apply(self, U, mu=None):
        Ub = U._blocks if self._block_source else [U]
        rs = self.range.subspaces if self._block_range else [self.range]
        nr, ns = self._blocks.shape
        out = []
        for i in range(nr):
            acc = None
            for j in range(ns):
                b = self._blocks[i, j]
                if b is None:
                    continue
                W = b.apply(Ub[j], mu=mu)
                if acc is None:
                    acc = W
                else:
                    acc.axpy(1.0, W)
            out.append(acc if acc is not None else rs[i].zeros(len(U)))
        return BlockVectorArray(out, self.range) if self._block_range else out[0]


class BlockDiagonalOperator(BlockOperator):
    def __init__(self, blocks, name=None, range_id=None, source_id=None):
        n = len(blocks)
        arr = np.full((n, n), None, dtype=object)
        for i, b in enumerate(blocks):
            arr[i, i] = b
        super().__init__(arr, name=name, range_id=range_id, source_id=source_id)


class BlockProjectionOperator(Operator):
    """Picks component ``index`` of a block array (fork-only; reference ``discretize...:696,704,714``)."""

    def __init__(self, block_space, index):
        self.source, self.index = block_space, int(index)
        self.range = block_space.subspaces[self.index]

    def apply(self, U, mu=None):
        return U._blocks[self.index].copy()

    @property
    def T(self):
        return BlockEmbeddingOperator(self.source, self.index)


class BlockEmbeddingOperator(Operator):
    def __init__(self, block_space, index):
        self.range, self.index = block_space, int(index)
        self.source = block_space.subspaces[self.index]

    def apply(self, U, mu=None):
        R = self.range.zeros(len(U))
        R._blocks[self.index] = U.copy()
        return R

    @property
    def T(self):
        return BlockProjectionOperator(self.range, self.index)


class BlockRowOperator(BlockOperator):
    """1 x S block operator with a non-block range (fork-only; reference ``discretize...:705,715``)."""
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


class _NeighborhoodMapOperator(Operator):
    """Common part of the Oswald-interpolation-error and flux-reconstruction operators: a linear map from
    ``domain_i`` into the block space over the neighbourhood of ``i``, one sparse matrix per component.

    The reference applies these vector by vector with C++ grid walks (``discretize...:83-122,148-176``); both are
    linear (``linear = True`` at ``:74,127``), so here each component is a CSR matrix and ``apply`` is one SpMM per
    component."""

    def __init__(self, subdomain, source_space, range_space, neighborhood, components, name=None):
        self.subdomain, self.neighborhood = int(subdomain), list(neighborhood)
        self.source, self.range = source_space, range_space
        self.components = [c if isinstance(c, DeviceCsr) else DeviceCsr(c) for c in components]
        self.name = name
        assert len(self.components) == len(self.range.subspaces)

    def apply(self, U, mu=None):
        return BlockVectorArray([spmm_once(c, U, s) for c, s in zip(self.components, self.range.subspaces)], self.range)


class OswaldInterpolationErrorOperator(_NeighborhoodMapOperator):
    """reference ``discretize_elliptic_block_swipdg.py:72-122``; range ``OI_i = (+)_{k in N(i)} domain_k`` (``:79-81``)."""

    def __init__(self, subdomain, solution_space, neighborhood, components):
        rng = BlockVectorSpace([solution_space.subspaces[ii] for ii in neighborhood], 'OI_{}'.format(subdomain))
        super().__init__(subdomain, solution_space.subspaces[subdomain], rng, neighborhood, components,
                         name='oswald_interpolation_error_{}'.format(subdomain))


class FluxReconstructionOperator(_NeighborhoodMapOperator):
    """reference ``discretize_elliptic_block_swipdg.py:125-176``; range ``RT_i = (+)_{k in N(i)} LOCALRT_k`` (``:140-146``)."""

    def __init__(self, subdomain, solution_space, neighborhood, rt_dims, components):
        rng = BlockVectorSpace([GpuVectorSpace(rt_dims[ii], 'LOCALRT_' + str(ii)) for ii in neighborhood],
                               'RT_{}'.format(subdomain))
        super().__init__(subdomain, solution_space.subspaces[subdomain], rng, neighborhood, components,
                         name='flux_reconstruction_{}'.format(subdomain))
