"""cProfile of one warm enrichment step at C2 (where does the host time go?)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pylrbms_b200 import LRBMSReductor, discretize
from pylrbms_b200.online_enrichment import AdaptiveEnrichment
from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases

data = assemble_block_swipdg((8, 8), 32)
S = data.num_subdomains
bases = make_local_bases(data, 20, seed=1002)
d, _ = discretize(data)
red = LRBMSReductor(d, bases={'domain_%d' % i: bases[i] for i in range(S)},
                    products=[d.operators['local_energy_dg_product_%d' % i] for i in range(S)])
red.incremental = True
rd = red.reduce()
ae = AdaptiveEnrichment(None, d, d.solution_space, red, rd, 1e-12, 0.33, 4)
for step, mu in enumerate((0.3, 0.7, 0.5, 0.9)):
    pr = cProfile.Profile()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pr.enable()
    ae.solve(mu, enrichment_steps=1)
    torch.cuda.synchronize()
    pr.disable()
    print('step %d (mu %.1f): %.3f s' % (step, mu, time.perf_counter() - t0), getattr(red.last_plan, 'timings', None))
    if step in (1, 3):
        st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(45); print(st.getvalue()[:9000])
