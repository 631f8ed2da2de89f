"""Generate the golden vectors in this directory from the CPU oracle (``oracle/``).

    python tests/golden/make_golden.py

The reference's third-party layers (pyMOR fork, dune-gdt) cannot be imported or built in this environment and its own
tests pin no result of the hot path, so these vectors are *oracle* outputs on seeded synthetic
inputs: they freeze the oracle (``tests/test_oracle_golden.py`` re-derives them on the CPU) and give the CUDA path a
fixed target (``tests/test_gpu_golden.py``).  The companion ``make_reference_golden.py`` produces the same quantities by
executing the reference's own ``estimators.py`` / ``reductor.py`` / ``online_enrichment.py`` (``reference_run__*.npz``).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    # name: (num_subdomains, cells_per_subdomain, basis sizes, basis seed, mu_bar, mu_hat)
    'os2015_2x2_N5': ((2, 2), 4, 5, 1001, 1.0, 1.0),
    'os2015_3x2_ragged': ((3, 2), 4, [3, 7, 4, 6, 5, 8], 1002, 0.6, 0.3),
    # 3D structure of BASELINE.json config C4 (six face neighbours) on seeded synthetic operators
    # (pylrbms_b200/synthetic_fixture.py): (grid, cells per subdomain and direction), ragged basis sizes
    'synthetic3d_3x2x2_ragged': ((3, 2, 2), (2, 2, 2), [5, 6, 7, 4, 8, 6, 5, 7, 6, 5, 4, 6], 1004, 0.7, 0.4),
}
MUS = np.array([0.1, 0.25, 0.5, 0.8, 1.0])


def input_digest(data, bases):
    h = hashlib.sha256()
    for q in range(data.Q):
        for key in sorted(data.lhs[q]):
            M = data.lhs[q][key]
            h.update(np.ascontiguousarray(M.indptr, dtype=np.int32).tobytes())
            h.update(np.ascontiguousarray(M.indices, dtype=np.int32).tobytes())
            h.update(np.round(M.data, 10).tobytes())
    for b in bases:
        h.update(np.round(b, 10).tobytes())
    return h.hexdigest()


def build_case(name):
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases, os2015_problem
    num_subdomains, cells, sizes, seed, mu_bar, mu_hat = CASES[name]
    if isinstance(cells, tuple):
        from pylrbms_b200.synthetic_fixture import make_random_local_bases, synthetic_block_operators
        data = synthetic_block_operators(num_subdomains, cells, seed=seed, problem=os2015_problem(mu_bar=mu_bar, mu_hat=mu_hat))
        return data, make_random_local_bases(data, sizes, seed=seed)
    data = assemble_block_swipdg(num_subdomains, cells, problem=os2015_problem(mu_bar=mu_bar, mu_hat=mu_hat))
    bases = make_local_bases(data, sizes, seed=seed)
    return data, bases


def oracle_outputs(data, bases):
    from oracle import lrbms_oracle as O
    from oracle.pymor_like import LincombOperator
    S = data.num_subdomains
    red = O.LRBMSReductor(O.build_discretization(data), bases={'domain_%d' % i: bases[i] for i in range(S)})
    rd = red.reduce()
    out = {'mus': MUS, 'block_dims': np.array(rd.block_dims)}
    for name, op in list(rd.operators.items()) + [('product_' + k, v) for k, v in rd.products.items()]:
        terms = op.operators if isinstance(op, LincombOperator) else [op]
        for q, t in enumerate(terms):
            M = t.matrix if hasattr(t, 'matrix') else t._array.data
            out['red__{}__{}'.format(name, q)] = np.asarray(M)
    U, eta, parts, ind = [], [], [], []
    for mu in MUS:
        u = rd.solve(mu)
        e, p, i_ = rd.estimate(u, mu, decompose=True)
        U.append(u.data[0]); eta.append(e); parts.append(np.stack([x[:, 0] for x in p])); ind.append(i_[:, 0])
    out.update(U=np.array(U), eta=np.array(eta), parts=np.array(parts), indicators=np.array(ind))
    return out


def main():
    for name in CASES:
        data, bases = build_case(name)
        out = oracle_outputs(data, bases)
        out['input_sha256'] = np.array(input_digest(data, bases))
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, '->', path, os.path.getsize(path), 'bytes', out['input_sha256'])


if __name__ == '__main__':
    main()
