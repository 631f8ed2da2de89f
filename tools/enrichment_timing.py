"""Wall time of ONE enrichment step at the C2 size (8x8 subdomains, n_i = 6144, N = 20): solve + estimate + Doerfler marking
+ local corrector solves (restarted CG on the neighbourhood systems) + Gram-Schmidt extension + incremental re-projection."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pylrbms_b200 import LRBMSReductor, discretize
from pylrbms_b200.online_enrichment import AdaptiveEnrichment
from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases

sx = int(os.environ.get('SUBDOMAINS', '8'))
data = assemble_block_swipdg((sx, sx), 32)
S = data.num_subdomains
bases = make_local_bases(data, 20, seed=1002)
d, _ = discretize(data)
red = LRBMSReductor(d, bases={'domain_%d' % i: bases[i] for i in range(S)},
                    products=[d.operators['local_energy_dg_product_%d' % i] for i in range(S)])
red.incremental = True
rd = red.reduce()
ae = AdaptiveEnrichment(None, d, d.solution_space, red, rd, 1e-12, 0.33, 4)
for mu in (0.3, 0.7, 0.5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    info_log = []
    # instrument the parts
    t_corr = [0.0]; n_corr = [0]; it_corr = [0]
    orig = d.solve_for_local_correction
    def timed(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = orig(*a, **k)
        torch.cuda.synchronize(); t_corr[0] += time.perf_counter() - t; n_corr[0] += 1
        it_corr[0] += d.last_local_correction_info['iterations']
        return out
    d.solve_for_local_correction = timed
    # Gram-Schmidt extension and re-projection, timed the same way
    t_gs = [0.0]; n_gs = [0]; t_red = [0.0]
    orig_gs, orig_reduce = red.extend_basis_local, red.reduce
    def timed_gs(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = orig_gs(*a, **k)
        torch.cuda.synchronize(); t_gs[0] += time.perf_counter() - t; n_gs[0] += 1
        return out
    def timed_reduce(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = orig_reduce(*a, **k)
        torch.cuda.synchronize(); t_red[0] += time.perf_counter() - t
        return out
    red.extend_basis_local, red.reduce = timed_gs, timed_reduce
    U, rd, _ = ae.solve(mu, enrichment_steps=1, callback=lambda rd_, U_, mu_, info: info_log.append(info))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    d.solve_for_local_correction = orig
    del red.extend_basis_local, red.reduce
    print('          of which: %d Gram-Schmidt extensions %.4f s; incremental re-projection (reduce()) %.3f s; rest (two reduced '
          'solves + estimates with their plan creation, marking, neighbourhood assembly) %.3f s'
          % (n_gs[0], t_gs[0], t_red[0], (t1 - t0) - t_corr[0] - t_gs[0] - t_red[0]))
    print('mu %.2f: one enrichment step %.3f s wall; %d corrector solves %.3f s (%d CG iterations, neighbourhood systems of up to '
          '%d dofs); eta %.4e -> %.4e; n_red %d -> %d' % (mu, t1 - t0, n_corr[0], t_corr[0], it_corr[0],
          d.last_local_correction_info['size'], info_log[0]['eta'], info_log[-1]['eta'], info_log[0]['global RB size'],
          info_log[-1]['global RB size']))
# the batched loop: 64 parameters per pass
mus = np.linspace(0.1, 1.0, 64)
torch.cuda.synchronize(); t0 = time.perf_counter()
log = []
ae.solve_batch(mus, enrichment_steps=1, callback=lambda rd_, U_, mu_, info: log.append(info))
torch.cuda.synchronize()
print('solve_batch, 64 parameters, one enrichment pass: %.3f s wall; eta_max %.4e -> %.4e' % (time.perf_counter() - t0, log[0]['eta_max'], log[-1]['eta_max']))
