"""Developer aid: full vs incremental re-projection after an enrichment on the bench workload (C2 by default).

    python tools/incremental_timing.py [bench.py options]
"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from pylrbms_b200 import build; build.build()
from pylrbms_b200 import LRBMSReductor, discretize

a = bench.parse_args()
data, bases = bench.make_inputs(a)
d, _ = discretize(data)
S = data.num_subdomains
products = [d.operators['local_energy_dg_product_%d' % k] for k in range(S)]
rng = np.random.default_rng(0)


def timed_reduce(red):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    planner_before = red.last_plan
    rd = red.reduce()
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    p = red.last_plan
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    return rd, wall, p


for rep, mode in enumerate(('full', 'incremental', 'full', 'incremental')):
    rng = np.random.default_rng(0)
    red = LRBMSReductor(d, bases=bases, products=products)
    red.incremental = (mode == 'incremental')
    rd, wall0, p0 = timed_reduce(red)
    out = ['%-11s first reduce %.3f s (%d descriptors)' % (mode, wall0, p0.n_project_descs)]
    for n_enriched in (1, 6, 64):
        for k in rng.choice(S, n_enriched, replace=False):
            red.extend_basis_local(d.solution_space.subspaces[k].from_data(rng.standard_normal((1, int(data.n[k])))))
        rd, wall, p = timed_reduce(red)
        st = p.stats()
        out.append('  +1 vector on %2d subdomains: reduce %.3f s wall, %5d descriptors (%d incremental jobs), %.2f GFLOP planned'
                   % (n_enriched, wall, p.n_project_descs, p.n_incremental_jobs, st['flops'] / 1e9))
    if rep >= 2:          # the first pass of each mode only loads the kernel variants it needs (lazy module loading)
        print('\n'.join(out))
