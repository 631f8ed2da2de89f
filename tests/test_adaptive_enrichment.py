"""Online adaptive enrichment: Doerfler marking, the local corrector solve and the
enrichment loop of reference ``online_enrichment.py`` / ``reductor.py:75-78`` / ``discretize...:227-316``.

CPU tests pin the marking rule against the line-by-line restatement in the oracle and a brute-force definition; GPU
tests compare the PCG neighbourhood solve with a sparse direct solve and the whole enrichment loop with the oracle's.
"""
import numpy as np
import pytest

ETA_TOL = 1e-10


def _brute_force_doerfler(indicators, theta):
    sq = np.asarray(indicators, dtype=float) ** 2
    order = sorted(range(len(sq)), key=lambda i: -sq[i])           # stable: ties keep index order
    total = np.sum([sq[i] for i in order])
    for k in range(len(order)):
        if np.sum([sq[i] for i in order[:k + 1]]) > theta * total:      # same prefix sums as the reference (numpy pairwise)
            return order[:k + 1]
    return order


@pytest.mark.parametrize('theta', [0.1, 0.33, 0.8, 1.0])
def test_doerfler_marking_matches_reference_restatement(theta):
    from oracle import lrbms_oracle as O
    from pylrbms_b200.online_enrichment import doerfler_marking
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 64):
        ind = rng.uniform(0.0, 1.0, n)
        ind[rng.integers(0, n)] = 0.0
        got = doerfler_marking(ind, theta)
        assert got == [int(i) for i in O.doerfler_marking(list(ind), theta)]
        assert got == _brute_force_doerfler(ind, theta)
    # ties and the "nothing exceeds" branch (theta = 1: the strict inequality never holds -> everything is marked)
    assert doerfler_marking([1.0, 1.0, 1.0], 1.0) == [0, 1, 2]
    assert doerfler_marking([2.0, 1.0, 2.0], 0.4) == [0]


def test_oracle_enrichment_loop():
    from oracle import lrbms_oracle as O
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg
    data = assemble_block_swipdg((2, 2), 4)
    d = O.build_discretization(data)
    products = [d.operators['local_energy_dg_product_%d' % i] for i in range(data.num_subdomains)]
    red = O.LRBMSReductor(d, products=products, order=0)
    rd = red.reduce()
    ae = O.AdaptiveEnrichment(None, d, d.solution_space, red, rd, 1e-12, 0.5, 2)
    # one enrichment step per parameter: the corrector problem does not depend on the current solution (its boundary
    # functional is commented out in the reference), so enriching one subdomain twice for the same parameter adds nothing
    # and raises ExtensionError there as here
    for mu in (0.5, 0.8, 0.2):
        etas = []
        ae.solve(mu, enrichment_steps=1, callback=lambda rd_, U, mu_, info: etas.append(info['eta']))
        assert len(etas) == 2 and etas[1] != etas[0] and np.isfinite(etas).all()
    # (enriching only the marked subdomains need not lower eta: the jumps to un-enriched neighbours grow)
    assert ae.rd.solution_space.dim > rd.solution_space.dim
    with pytest.raises(O.ExtensionError):
        ae.solve(0.2, enrichment_steps=3)


@pytest.mark.gpu
def test_pcg_matches_sparse_direct_solve(handle):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    from pylrbms_b200.kernels import DeviceCsr, pcg_solve
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg
    data = assemble_block_swipdg((2, 2), 8)
    nb = data.neighborhoods[0]
    A = sp.bmat([[data.lhs[0].get((k, l)) for l in nb] for k in nb], format='csr') + \
        0.3 * sp.bmat([[data.lhs[1].get((k, l)) for l in nb] for k in nb], format='csr')
    b = np.concatenate([data.rhs[k] for k in nb])
    x_ref = spla.spsolve(A.tocsc(), b)
    x, iters, relres = pcg_solve(DeviceCsr(A), torch.from_numpy(b).cuda(), rtol=1e-13)
    assert 0 < iters < 10 * A.shape[0] and relres <= 1e-13
    assert np.abs(x.cpu().numpy() - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    # warm start from the solution: nothing left to do
    x2, iters2, _ = pcg_solve(DeviceCsr(A), torch.from_numpy(b).cuda(), x0=x, rtol=1e-12)
    assert iters2 == 0 and torch.equal(x2, x)
    # the cooperative kernel (25 iterations per launch, grid-wide barriers) and three launches per iteration run the
    # same recurrences: same iteration count, same solution to rounding, both reproducible bit for bit
    x_again, iters_again, _ = pcg_solve(DeviceCsr(A), torch.from_numpy(b).cuda(), rtol=1e-13)
    assert iters_again == iters and torch.equal(x_again, x)
    handle.check(handle.lib.lrbms_set_option(handle.h, 2, 1))          # LRBMS_OPT_PCG_MULTI_LAUNCH
    try:
        x3, iters3, relres3 = pcg_solve(DeviceCsr(A), torch.from_numpy(b).cuda(), rtol=1e-13)
    finally:
        handle.check(handle.lib.lrbms_set_option(handle.h, 2, 0))
    assert relres3 <= 1e-13 and abs(iters3 - iters) <= 25
    assert np.abs(x3.cpu().numpy() - x.cpu().numpy()).max() <= 1e-11 * np.abs(x_ref).max()
    # an indefinite matrix is reported, not silently "solved"
    from pylrbms_b200._lib import LrbmsError
    with pytest.raises(LrbmsError):
        pcg_solve(DeviceCsr(-A), torch.from_numpy(b).cuda())


@pytest.mark.gpu
@pytest.mark.parametrize('num_subdomains,cells', [((2, 2), 4), ((3, 3), 4)])
def test_adaptive_enrichment_matches_oracle(handle, num_subdomains, cells):
    from oracle import lrbms_oracle as O
    from pylrbms_b200 import LRBMSReductor, discretize
    from pylrbms_b200.online_enrichment import AdaptiveEnrichment
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, spe10_like_problem
    # a field without symmetries: on the symmetric OS2015 example mirror-image subdomains have equal indicators and the
    # Doerfler marking then hangs on the last bit of the estimator
    data = assemble_block_swipdg(num_subdomains, cells, problem=spe10_like_problem(seed=7, contrast=50.0))
    S = data.num_subdomains
    # oracle
    d_ref = O.build_discretization(data)
    red_ref = O.LRBMSReductor(d_ref, products=[d_ref.operators['local_energy_dg_product_%d' % i] for i in range(S)], order=0)
    log_ref = []
    mus = (0.4, 0.9, 0.15)
    ae_ref = O.AdaptiveEnrichment(None, d_ref, d_ref.solution_space, red_ref, red_ref.reduce(), 1e-12, 0.5, 2)
    for mu in mus:          # one enrichment step per parameter (see test_oracle_enrichment_loop)
        U_ref, rd_ref, _ = ae_ref.solve(mu, enrichment_steps=1, callback=lambda rd_, U, mu_, info: log_ref.append(info))
    # CUDA path
    d, _ = discretize(data)
    red = LRBMSReductor(d, products=[d.operators['local_energy_dg_product_%d' % i] for i in range(S)], order=0)
    log = []
    ae = AdaptiveEnrichment(None, d, d.solution_space, red, red.reduce(), 1e-12, 0.5, 2)
    for mu in mus:
        U, rd, red_out = ae.solve(mu, enrichment_steps=1, callback=lambda rd_, U_, mu_, info: log.append(info))
    assert red_out is red
    assert len(log) == len(log_ref) == 6
    for a, b in zip(log, log_ref):
        assert a['local_problem_solves'] == b['local_problem_solves']
        assert a['global RB size'] == b['global RB size']
        # the corrector comes from restarted CG (true residual 1e-13 of ||b||, i.e. the attainable accuracy of the oracle's
        # sparse direct solve) and is then orthonormalised and projected: 1e-10 tolerance like the rest of the path plus
        # the conditioning of the corrector problem (both solvers carry eps * cond(A_nbh) ~ 1e-11 of their own)
        worst_eta = max(locals().get('worst_eta', 0.0), abs(a['eta'] - b['eta']) / abs(b['eta']))
        assert abs(a['eta'] - b['eta']) <= ETA_TOL * abs(b['eta'])
    assert rd.block_dims == rd_ref.block_dims
    from pylrbms_b200 import ExtensionError
    with pytest.raises(ExtensionError):          # same parameter again: nothing new to add (reference behaviour)
        ae.solve(mus[-1], enrichment_steps=3)
    # the enriched reduced solutions describe the same fine-scale function
    u_fine = red.reconstruct(U).to_numpy()[0]
    u_fine_ref = np.concatenate([b.data[0] for b in red_ref.reconstruct(U_ref)._blocks])
    print('enrichment parity: worst eta rel diff', worst_eta, 'fine-scale max rel diff',
          np.abs(u_fine - u_fine_ref).max() / np.abs(u_fine_ref).max(), d.last_local_correction_info)
    assert np.abs(u_fine - u_fine_ref).max() <= 10 * ETA_TOL * np.abs(u_fine_ref).max()
    assert d.last_local_correction_info['relative_residual'] <= 1e-13


def _sequence_fixture(name):
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, 'golden'))
    import make_reference_golden as M
    gold = np.load(os.path.join(here, 'golden', 'reference_run__' + name + '.npz'))
    data, mus = M.build_sequence_case(name)
    assert list(mus) == list(gold['mus'])
    return gold, data, mus


@pytest.mark.parametrize('name', ['enrichment_spe10like_2x2', 'enrichment_spe10like_3x3'])
def test_oracle_enrichment_matches_reference_run(name):
    """The oracle's enrichment loop against the committed run of the REFERENCE's ``AdaptiveEnrichment`` / ``enrich_local``
    (``oracle/reference_run.py``; order-0 shape-function bases, energy products, one step per parameter)."""
    from oracle import lrbms_oracle as O
    gold, data, mus = _sequence_fixture(name)
    S = data.num_subdomains
    d = O.build_discretization(data)
    red = O.LRBMSReductor(d, products=[d.operators['local_energy_dg_product_%d' % i] for i in range(S)], order=0)
    theta, max_age, target = gold['args']
    ae = O.AdaptiveEnrichment(None, d, d.solution_space, red, red.reduce(), float(target), float(theta), int(max_age))
    log = []
    for mu in mus:
        U, rd, _ = ae.solve(mu, enrichment_steps=1, callback=lambda rd_, U_, mu_, info: log.append(info))
    assert [l['global RB size'] for l in log] == list(gold['rb_size'])
    assert [l['local_problem_solves'] for l in log] == list(gold['local_problem_solves'])
    assert np.abs(np.array([l['eta'] for l in log]) - gold['eta']).max() <= 1e-12 * np.abs(gold['eta']).max()
    assert list(rd.block_dims) == list(gold['block_dims'])
    u_fine = np.concatenate([b.data[0] for b in red.reconstruct(U)._blocks])
    assert np.abs(u_fine - gold['u_fine']).max() <= 1e-12 * np.abs(gold['u_fine']).max()


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['enrichment_spe10like_2x2', 'enrichment_spe10like_3x3'])
def test_adaptive_enrichment_matches_reference_run(handle, name):
    """The CUDA enrichment loop (``AdaptiveEnrichment.solve`` -> ``enrich_local`` -> ``lrbms_pcg_solve`` -> Gram-Schmidt ->
    incremental ``reduce``) against the committed run of the reference's own loop on the same inputs."""
    from pylrbms_b200 import LRBMSReductor, discretize
    from pylrbms_b200.online_enrichment import AdaptiveEnrichment
    gold, data, mus = _sequence_fixture(name)
    S = data.num_subdomains
    d, _ = discretize(data)
    red = LRBMSReductor(d, products=[d.operators['local_energy_dg_product_%d' % i] for i in range(S)], order=0)
    theta, max_age, target = gold['args']
    ae = AdaptiveEnrichment(None, d, d.solution_space, red, red.reduce(), float(target), float(theta), int(max_age))
    log = []
    for mu in mus:
        U, rd, _ = ae.solve(mu, enrichment_steps=1, callback=lambda rd_, U_, mu_, info: log.append(info))
    assert [l['global RB size'] for l in log] == list(gold['rb_size'])
    assert [l['local_problem_solves'] for l in log] == list(gold['local_problem_solves'])
    eta = np.array([l['eta'] for l in log])
    print('enrichment vs reference run: worst eta rel diff', np.abs(eta / gold['eta'] - 1).max())
    assert np.abs(eta - gold['eta']).max() <= ETA_TOL * np.abs(gold['eta']).max()
    assert list(rd.block_dims) == list(gold['block_dims'])
    u_fine = red.reconstruct(U).to_numpy()[0]
    assert np.abs(u_fine - gold['u_fine']).max() <= 10 * ETA_TOL * np.abs(gold['u_fine']).max()


@pytest.mark.gpu
def test_batched_enrichment(handle):
    """``AdaptiveEnrichment.solve_batch``: one sweep per pass over the whole parameter batch, enrichment from the worst
    parameter.  A batch of one must retrace ``solve`` exactly; on a real batch the largest estimate over the batch must be
    what every pass reports, the bases must grow, and the final model must reproduce a fresh sweep."""
    from pylrbms_b200 import LRBMSReductor, discretize
    from pylrbms_b200.online_enrichment import AdaptiveEnrichment
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, spe10_like_problem
    data = assemble_block_swipdg((3, 3), 4, problem=spe10_like_problem(seed=7, contrast=50.0))
    S = data.num_subdomains

    def fresh():
        d, _ = discretize(data)
        red = LRBMSReductor(d, products=[d.operators['local_energy_dg_product_%d' % i] for i in range(S)], order=0)
        return d, red, AdaptiveEnrichment(None, d, d.solution_space, red, red.reduce(), 1e-12, 0.5, 2)
    # batch of one == the reference-shaped loop
    _, _, ae1 = fresh()
    log1 = []
    ae1.solve(0.4, enrichment_steps=1, callback=lambda rd_, U, mu_, info: log1.append(info))
    _, _, ae2 = fresh()
    log2 = []
    ae2.solve_batch([0.4], enrichment_steps=1, callback=lambda rd_, U, mu_, info: log2.append(info))
    assert len(log1) == len(log2) == 2
    for a, b in zip(log1, log2):
        assert a['global RB size'] == b['global RB size'] and a['eta'] == b['eta'][0]
    # a real batch
    _, red, ae = fresh()
    mus = np.linspace(0.15, 0.9, 9)
    log = []
    U, eta, rd, red_out = ae.solve_batch(mus, enrichment_steps=2, callback=lambda rd_, U_, mu_, info: log.append(info))
    assert red_out is red and len(log) == 3 and len(U) == len(mus)
    assert all(info['eta_max'] == info['eta'].max() and info['argmax'] == int(np.argmax(info['eta'])) for info in log)
    assert log[-1]['global RB size'] > log[0]['global RB size']
    U2, eta2 = rd.sweep(mus)
    assert np.array_equal(eta2, eta) and np.array_equal(U2.data, U.data)
