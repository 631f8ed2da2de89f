"""CPU tests of the oracle: it reproduces the committed golden vectors, agrees with an independent direct-formula
restatement of SURVEY.md Appendix B, and keeps the reference quirks.

The golden vectors are oracle outputs (the reference cannot run here -- "parity unpinned", see
tests/golden/make_golden.py); this file guards them against drift and cross-checks the operator-algebra oracle
(oracle/pymor_like.py + oracle/lrbms_oracle.py) against plain dense formulas written down independently.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden  # noqa: E402


@pytest.mark.parametrize('name', sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    gold = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    data, bases = make_golden.build_case(name)
    assert make_golden.input_digest(data, bases) == str(gold['input_sha256']), 'fixture generator drifted'
    out = make_golden.oracle_outputs(data, bases)
    for key in gold.files:
        if key == 'input_sha256':
            continue
        ref, got = gold[key], np.asarray(out[key])
        assert got.shape == ref.shape, key
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(got - ref).max() <= 1e-12 * scale, key


def _direct_formulas(data, bases):
    """SURVEY.md Appendix B written out with dense NumPy, no operator classes."""
    S, Q = data.num_subdomains, data.Q
    N = [b.shape[0] for b in bases]
    V = [b.T for b in bases]                                  # n_i x N_i
    out = {}
    for q in range(Q):
        for (i, j), M in data.lhs[q].items():
            out[('operator', q, i, j)] = V[i].T @ (M @ V[j])
    for i in range(S):
        out[('rhs', 0, 0, i)] = (data.rhs[i] @ V[i])[None, :]
        out[('l2', 0, i, i)] = V[i].T @ (data.l2[i] @ V[i])
        nb = data.neighborhoods[i]
        W = {k: data.oi[(k, i)] @ V[k] for k in nb}                                        # component i of OI_k(V_k)
        R = {k: np.hstack([data.fr[q][(k, i)] @ V[k] for q in range(Q)]) for k in nb}      # q-major RT basis slice
        DR = {k: data.div[i] @ R[k] for k in nb}
        for k in nb:
            out[('r_fd_%d' % i, 0, 0, k)] = (data.rhs[i] @ DR[k])[None, :]
            for q in range(Q):
                out[('df_ab_%d' % i, q, i, k)] = V[i].T @ (data.ab[q][i] @ R[k])
            for k2 in nb:
                out[('nc_%d' % i, 0, k, k2)] = W[k].T @ (data.elliptic[i] @ W[k2])
                out[('r_dd_%d' % i, 0, k, k2)] = DR[k].T @ (data.l2[i] @ DR[k2])
                out[('df_bb_%d' % i, 0, k, k2)] = R[k].T @ (data.bb[i] @ R[k2])
        for q in range(Q):
            for q2 in range(Q):
                out[('df_aa_%d' % i, q * Q + q2, i, i)] = V[i].T @ (data.aa[q][q2][i] @ V[i])
    return out


def test_oracle_matches_direct_formulas():
    from oracle import lrbms_oracle as O
    data, bases = make_golden.build_case('os2015_3x2_ragged')
    S = data.num_subdomains
    red = O.LRBMSReductor(O.build_discretization(data), bases={'domain_%d' % i: bases[i] for i in range(S)})
    red.reduce()
    direct = _direct_formulas(data, bases)
    seen = 0
    for (name, q, i, j), B in direct.items():
        got = O.reduced_blocks(red, name, q if name.startswith(('operator', 'df_aa', 'df_ab', 'rhs')) else None)
        assert (i, j) in got, (name, q, i, j)
        scale = max(np.abs(B).max(), 1e-12)
        assert np.abs(got[(i, j)] - B).max() <= 1e-11 * max(scale, 1.0), (name, q, i, j)
        seen += 1
    assert seen == len(direct) and seen > 200


def test_estimator_formula_and_quirks():
    """eta assembled by hand from the reduced matrices equals the oracle's estimate; alpha looks at theta_0 only."""
    from oracle import lrbms_oracle as O
    data, bases = make_golden.build_case('os2015_3x2_ragged')
    S, Q = data.num_subdomains, data.Q
    N = [b.shape[0] for b in bases]
    off = np.concatenate([[0], np.cumsum(N)])
    direct = _direct_formulas(data, bases)
    for flag in (True, False):
        d = O.build_discretization(data, alpha_returns_first=flag)
        rd = O.LRBMSReductor(d, bases={'domain_%d' % i: bases[i] for i in range(S)}).reduce()
        mu = 0.35
        th = [1.0, mu]
        U = rd.solve(mu)
        u = [U.data[0, off[i]:off[i + 1]] for i in range(S)]
        eta, (nc, r, df), ind = rd.estimate(U, mu, decompose=True)
        tb = [1.0, float(data.mu_bar['diffusion'][0])]
        th_hat = [1.0, float(data.mu_hat['diffusion'][0])]
        rb = [a / b for a, b in zip(th, tb)]
        rh = [a / b for a, b in zip(th, th_hat)]
        a_bar, a_hat = (rb[0], rh[0]) if flag else (min(rb), min(rh))
        g_bar = max(rb)
        nc2, r2, df2 = np.zeros(S), np.zeros(S), np.zeros(S)
        for i in range(S):
            nb = data.neighborhoods[i]
            ur = {k: np.concatenate([th[q] * u[k] for q in range(Q)]) for k in nb}
            for k in nb:
                r2[i] -= 2 * float((direct[('r_fd_%d' % i, 0, 0, k)] @ ur[k])[0])
                for q in range(Q):
                    df2[i] += 2 * th[q] * u[i] @ direct[('df_ab_%d' % i, q, i, k)] @ ur[k]
                for k2 in nb:
                    nc2[i] += u[k] @ direct[('nc_%d' % i, 0, k, k2)] @ u[k2]
                    r2[i] += ur[k] @ direct[('r_dd_%d' % i, 0, k, k2)] @ ur[k2]
                    df2[i] += ur[k] @ direct[('df_bb_%d' % i, 0, k, k2)] @ ur[k2]
            for q in range(Q):
                for q2 in range(Q):
                    df2[i] += th[q] * th[q2] * u[i] @ direct[('df_aa_%d' % i, q * Q + q2, i, i)] @ u[i]
            r2[i] = (data.local_eta_rf_squared[i] + r2[i]) * (1 / np.pi ** 2) / data.min_diffusion_evs[i] * data.subdomain_diameters[i] ** 2
        eta2 = (np.sqrt(g_bar) * np.linalg.norm(nc2) + np.linalg.norm(r2 + df2) / np.sqrt(a_hat)) / np.sqrt(a_bar)
        assert np.abs(nc[:, 0] - nc2).max() <= 1e-10 * np.abs(nc2).max()
        assert np.abs(df[:, 0] - df2).max() <= 1e-10 * np.abs(df2).max()
        assert np.abs(r[:, 0] - r2).max() <= 1e-10 * max(np.abs(r2).max(), 1e-3)
        assert abs(eta - eta2) <= 1e-10 * eta2
        ind2 = (2 / a_bar) * (g_bar * nc2 ** 2 + (r2 + df2) ** 2 / a_hat)     # squared again (estimators.py:106-107)
        assert np.abs(ind[:, 0] - ind2).max() <= 1e-9 * np.abs(ind2).max()
    # mu_bar = 0.6, mu_hat = 0.3 make the two alpha variants differ
    assert min(rb) != rb[0] or min(rh) != rh[0]


def test_per_vector_mode_equals_vectorised():
    """``MatrixOperator.per_vector`` walks one ``mv`` per basis vector like the reference's ListVectorArray."""
    from oracle import lrbms_oracle as O
    from oracle.pymor_like import MatrixOperator
    data, bases = make_golden.build_case('os2015_2x2_N5')
    bd = {'domain_%d' % i: bases[i] for i in range(4)}
    a = O.LRBMSReductor(O.build_discretization(data), bases=bd).reduce().operator.operators[1].matrix
    MatrixOperator.per_vector = True
    try:
        b = O.LRBMSReductor(O.build_discretization(data), bases=bd).reduce().operator.operators[1].matrix
    finally:
        MatrixOperator.per_vector = False
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


def test_reduced_operator_is_spd_and_solution_satisfies_system():
    from oracle import lrbms_oracle as O
    data, bases = make_golden.build_case('os2015_2x2_N5')
    rd = O.LRBMSReductor(O.build_discretization(data), bases={'domain_%d' % i: bases[i] for i in range(4)}).reduce()
    for mu in (0.1, 1.0):
        A = rd.operator.assemble(rd.parse_parameter(mu)).matrix
        assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()
        assert np.linalg.eigvalsh(A).min() > 0
        U = rd.solve(mu)
        f = rd.rhs.as_source_array(rd.parse_parameter(mu)).data[0]
        assert np.abs(A @ U.data[0] - f).max() <= 1e-12 * np.abs(f).max()


def test_parity_checker_scales_bound_the_forms():
    """``oracle/parity.py``: the summed-magnitude scales ``S`` the parity criterion uses must bound the estimator parts they
    belong to (|x^T M y| <= |x|^T |M| |y| term by term), and comparing the oracle with itself must give exact zeros."""
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases
    from oracle import lrbms_oracle as O
    from oracle.parity import ULPS, reference_online
    data = assemble_block_swipdg((3, 2), 4)
    bases = make_local_bases(data, [3, 5, 4, 6, 5, 4], seed=2)
    rd_ref = O.LRBMSReductor(O.build_discretization(data), bases={'domain_%d' % i: bases[i] for i in range(6)}).reduce()
    mus = [0.15, 0.6, 1.0]
    U, eta, parts, ind, As, S, consts = reference_online(rd_ref, mus)
    assert S.shape == parts.shape == (3, 6, 3)
    assert np.all(S >= np.abs(parts) * (1 - 1e-12))
    assert np.all(consts > 0) and 0 < ULPS < 1e-3
    # the scale of eta bounds eta itself
    a_bar, g_bar, a_hat = consts[:, 0], consts[:, 1], consts[:, 2]
    s_eta = (np.sqrt(g_bar) * np.linalg.norm(S[0], axis=0) + np.linalg.norm(S[1] + S[2], axis=0) / np.sqrt(a_hat)) / np.sqrt(a_bar)
    assert np.all(s_eta >= eta * (1 - 1e-12))
