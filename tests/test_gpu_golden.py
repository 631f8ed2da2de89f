"""The CUDA path against the committed golden vectors on seeded inputs: ``tests/golden/<case>.npz`` (oracle outputs) and
``tests/golden/reference_run__<case>.npz`` (outputs of the reference's own estimators.py / reductor.py, executed unmodified on
NumPy stand-ins for their third-party layers -- ``oracle/reference_run.py``)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden  # noqa: E402

RTOL = 1e-10


@pytest.mark.parametrize('source', ['', 'reference_run__'])
@pytest.mark.parametrize('name', sorted(make_golden.CASES))
def test_cuda_path_reproduces_golden(handle, name, source):
    from pylrbms_b200 import LRBMSReductor, discretize
    from pylrbms_b200.operators import LincombOperator
    gold = np.load(os.path.join(HERE, 'golden', source + name + '.npz'))
    data, bases = make_golden.build_case(name)
    assert make_golden.input_digest(data, bases) == str(gold['input_sha256'])
    S = data.num_subdomains
    red = LRBMSReductor(discretize(data)[0], bases={'domain_%d' % i: bases[i] for i in range(S)})
    rd = red.reduce()
    assert list(gold['block_dims']) == rd.block_dims
    checked = 0
    for key in gold.files:
        if not key.startswith('red__'):
            continue
        _, opname, q = key.split('__')
        op = rd.products[opname[len('product_'):]] if opname.startswith('product_') else rd.operators[opname]
        term = op.operators[int(q)] if isinstance(op, LincombOperator) else op
        ref = gold[key]
        got = term.to_dense()
        assert got.shape == ref.shape, key
        assert np.abs(got - ref).max() <= RTOL * max(np.abs(ref).max(), 1e-300), key
        checked += 1
    assert checked >= 10 * S
    for key in gold.files:                                     # Oswald / flux-reconstruction image bases (reference-run fixtures)
        if key.startswith('basis__'):
            a, b = red.bases[key[len('basis__'):]].to_numpy(), gold[key]
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-13 * max(1.0, np.abs(b).max()), key
    mus = gold['mus']
    U, eta, parts, ind = rd.sweep(mus, decompose=True)
    for k in range(len(mus)):                                  # energy norm of the assembled reduced operator, 1e-10
        A = sum(c * o.to_dense() for c, o in zip(rd.thetas([mus[k]])[0], rd.operator.operators))
        e = U.data[k] - gold['U'][k]
        assert np.sqrt(e @ A @ e) <= RTOL * np.sqrt(gold['U'][k] @ A @ gold['U'][k])
    assert np.abs(eta - gold['eta']).max() <= RTOL * np.abs(gold['eta']).max()
    gp = gold['parts']                                          # (n_mu, 3, S)
    r_floor = np.abs(rd.estimator.local_eta_rf_squared * rd.estimator.r_scale()).max()
    for kind in range(3):
        scale = max(np.abs(gp[:, kind]).max(), r_floor if kind == 1 else 0.0)
        assert np.abs(parts[kind].T - gp[:, kind]).max() <= RTOL * scale
    assert np.abs(ind.T - gold['indicators']).max() <= 10 * RTOL * np.abs(gold['indicators']).max()
    if 'fine_eta' in gold.files:
        # fine-scale estimator (generic operator chain on the GPU operators) on the reconstructed solutions, against the run of
        # the reference's estimator on the fine-scale operators; cancelling residual terms -> scale of the largest part
        d = red.d
        for k, mu in enumerate(mus):
            e, p, i_ = d.estimate(red.reconstruct(rd.solve(mu)), mu, decompose=True)
            assert abs(e - gold['fine_eta'][k]) <= 10 * RTOL * abs(gold['fine_eta'][k]), (k, e, gold['fine_eta'][k])
            got = np.stack([np.asarray(x).reshape(S, -1)[:, 0] for x in p])
            scale = max(np.abs(gold['fine_parts'][k]).max(), r_floor)
            assert np.abs(got - gold['fine_parts'][k]).max() <= 10 * RTOL * scale
