"""Where does the host time of LRBMSReductor.reduce() go? (planning is Python; the kernels take a few ms)"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pylrbms_b200 import LRBMSReductor, discretize
a = bench.parse_args()
data, bases = bench.make_inputs(a)
t = time.perf_counter(); d, _ = discretize(data); torch.cuda.synchronize(); print('discretize (upload)', time.perf_counter() - t)
red = LRBMSReductor(d, bases=bases)
for k in range(2):
    t = time.perf_counter(); rd = red.reduce(); torch.cuda.synchronize(); print('reduce #%d' % k, time.perf_counter() - t)
pr = cProfile.Profile(); pr.enable(); rd = red.reduce(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
t = time.perf_counter(); rd.online_plan; torch.cuda.synchronize(); print('online plan', time.perf_counter() - t)
