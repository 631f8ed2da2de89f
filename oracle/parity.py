"""Parity checker: CUDA reduced model (``pylrbms_b200``) against the oracle's reduced model on the same inputs.

TEST INFRASTRUCTURE (like everything under ``oracle/``): imported by ``tests/``, ``__graft_entry__.smoke()`` and the
parity gate of ``bench.py`` only -- never by the product package.  The functions take the product objects as arguments
and use nothing but their public API (``to_dense``, ``sweep``, ``sweep_into``), so the comparison goes through the same
calls a user makes.

What is compared (tolerance 1e-10 relative, BASELINE.json north_star; scales per SURVEY.md section 7 "hard parts"):

* every reduced operator and product the reference's ``reductor.reduce()`` returns (``reductor.py:70`` ->
  ``GenericRBSystemReductor._reduce``), as the unblocked matrix the reference stores (``unblock``, ``reductor.py:46,66``),
  max-norm relative to the operator's largest entry;
* ``u(mu)`` in the energy norm of the assembled reduced operator ``A(mu)``, relative to ``||u_ref||_A``;
* ``eta(mu)``, and its three parts: ``nc`` and ``df`` relative to their largest value over the subdomains, the cancelling
  residual ``r = (||f||^2 - 2 r_fd + r_dd) * scale`` relative to ``||f||^2 * scale`` (``estimators.py:72-76, 88-91``).
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-10


def _dense_ref(op):
    """The unblocked matrix (or row vector) of an oracle reduced operator."""
    if hasattr(op, 'matrix'):
        return np.asarray(op.matrix)
    return np.asarray(op._array.data)


def compare_operators(rd, rd_ref, names=None):
    """``{name: rel err}`` over all reduced operators / products; affine components are compared one by one."""
    out = {}
    if names is None:
        names = list(rd_ref.operators) + ['product:' + k for k in rd_ref.products]
    for name in names:
        if name.startswith('product:'):
            got, ref = rd.products[name[8:]], rd_ref.products[name[8:]]
        else:
            got, ref = rd.operators[name], rd_ref.operators[name]
        gots = got.operators if hasattr(got, 'operators') else [got]
        refs = ref.operators if hasattr(ref, 'operators') else [ref]
        assert len(gots) == len(refs), '{}: {} vs {} affine components'.format(name, len(gots), len(refs))
        worst = 0.0
        for g, r in zip(gots, refs):
            B = _dense_ref(r)
            A = g.to_dense()
            assert A.shape == B.shape, '{}: shape {} vs {}'.format(name, A.shape, B.shape)
            scale = np.abs(B).max()
            if scale == 0.0:
                worst = max(worst, float(np.abs(A).max()))
            else:
                worst = max(worst, float(np.abs(A - B).max() / scale))
        out[name] = worst
    return out


def reference_online(rd_ref, mus):
    """What the reference does per parameter (``online_enrichment.py:72-74``): returns ``U (n_mu, n_red)``, ``eta``,
    ``parts (3, S, n_mu)``, ``indicators (S, n_mu)`` and the assembled operators."""
    U, eta, parts, ind, As = [], [], [], [], []
    for mu in mus:
        u = rd_ref.solve(mu)
        e, p, i = rd_ref.estimate(u, mu, decompose=True)
        U.append(u.data[0]); eta.append(e)
        parts.append(np.stack([np.asarray(x)[:, 0] for x in p])); ind.append(np.asarray(i)[:, 0])
        As.append(np.asarray(rd_ref.operator.assemble(rd_ref.parse_parameter(mu)).matrix))
    return np.array(U), np.array(eta), np.stack(parts, axis=2), np.stack(ind, axis=1), As


def compare_online(rd, rd_ref, mus, U=None, eta=None, parts=None, ind=None, ref=None):
    """Relative errors of the online results for ``mus``.  ``U, eta, parts, ind`` default to ``rd.sweep(mus, decompose=True)``;
    pass what another entry point (``sweep_into``) produced to check that one instead.  ``ref`` caches ``reference_online``."""
    if ref is None:
        ref = reference_online(rd_ref, mus)
    U_ref, eta_ref, parts_ref, ind_ref, As = ref
    if U is None:
        Ua, eta, p, ind = rd.sweep(mus, decompose=True)
        U, parts = np.asarray(Ua.data), np.stack(p)
    errs = {}
    en = 0.0
    for k, A in enumerate(As):
        e = U[k] - U_ref[k]
        en = max(en, float(np.sqrt(max(e @ A @ e, 0.0)) / np.sqrt(U_ref[k] @ A @ U_ref[k])))
    errs['u_energy'] = en
    if eta is not None:
        errs['eta'] = float(np.max(np.abs(np.asarray(eta) - eta_ref) / np.abs(eta_ref)))
    if parts is not None:
        est = rd.estimator
        r_scale = np.abs(np.asarray(est.local_eta_rf_squared) * np.asarray(est.r_scale()))
        for name, k, extra in (('nc', 0, 0.0), ('r', 1, float(r_scale.max())), ('df', 2, 0.0)):
            s = max(float(np.abs(parts_ref[k]).max()), extra)
            errs[name] = float(np.abs(parts[k] - parts_ref[k]).max() / s) if s > 0 else float(np.abs(parts[k]).max())
    if ind is not None:
        errs['indicators'] = float(np.abs(np.asarray(ind) - ind_ref).max() / np.abs(ind_ref).max())
    return errs


def assert_parity(rd, rd_ref, mus, rtol=RTOL, what='all'):
    """Raise AssertionError with the offending quantity; returns ``(max rel err, number of checked quantities)``."""
    worst, n = 0.0, 0
    if what in ('all', 'offline'):
        for name, err in compare_operators(rd, rd_ref).items():
            assert err <= rtol, 'reduced operator {}: rel err {:.3e} > {:.1e}'.format(name, err, rtol)
            worst, n = max(worst, err), n + 1
    if what in ('all', 'online'):
        for name, err in compare_online(rd, rd_ref, mus).items():
            lim = 10 * rtol if name == 'indicators' else rtol     # indicators square the parts (estimators.py:106-107)
            assert err <= lim, 'online {}: rel err {:.3e} > {:.1e}'.format(name, err, lim)
            worst, n = max(worst, err), n + len(mus)
    return worst, n
