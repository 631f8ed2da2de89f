"""Where does the wall time of the first / repeated LRBMSReductor.reduce() go?  (host planner vs plan creation vs kernels)"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pylrbms_b200 import LRBMSReductor, discretize
from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases

sx = int(os.environ.get('SUBDOMAINS', '8'))
data = assemble_block_swipdg((sx, sx), 32)
bases = make_local_bases(data, 20, seed=1002)
bd = {'domain_%d' % i: bases[i] for i in range(data.num_subdomains)}
t = time.perf_counter(); d, _ = discretize(data); torch.cuda.synchronize(); print('discretize s', time.perf_counter() - t)
red = LRBMSReductor(d, bases=bd)
for call in range(3):
    pr = cProfile.Profile()
    torch.cuda.synchronize(); t = time.perf_counter()
    pr.enable(); rd = red.reduce(); torch.cuda.synchronize(); pr.disable()
    print('reduce() call', call, 's', time.perf_counter() - t, red.last_plan.timings)
    if call in (0, 2):
        st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(28); print(st.getvalue()[:6000])
