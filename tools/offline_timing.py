"""Developer aid: time the offline projection plan of the bench workload (C2 by default) with CUDA events.

    python tools/offline_timing.py [bench.py options] ; SINGLE_STREAM=1 serialises the buckets on one stream
    (lrbms_set_option(LRBMS_OPT_SINGLE_STREAM), e.g. for ncu captures)
"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from pylrbms_b200 import build; build.build()
from pylrbms_b200 import LRBMSReductor, discretize

a = bench.parse_args()
data, bases = bench.make_inputs(a)
d, _ = discretize(data)
from pylrbms_b200._lib import Handle
_h = Handle.get()
_single = os.environ.get('SINGLE_STREAM', '0') not in ('', '0')
_h.check(_h.lib.lrbms_set_option(_h.h, 1, 1 if _single else 0))
reductor = LRBMSReductor(d, bases=bases)
rd = reductor.reduce()
torch.cuda.synchronize()
planner = reductor.last_plan
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device='cuda')
for _ in range(3):
    planner.run()
ta, tp = [], []
for _ in range(int(os.environ.get('REPS', '10'))):
    flush.fill_(1.0)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    planner.spmm_plans[0].run()            # stage 0: OI / FR images of the bases
    e1.record()
    for p in planner.spmm_plans[1:]:       # later SpMM stages are part of the projection
        p.run()
    planner.project_plan.run()
    e2.record(); e2.synchronize()
    ta.append(e0.elapsed_time(e2)); tp.append(e1.elapsed_time(e2))
pp = planner.project_plan
t = float(np.mean(tp)) * 1e-3
print('single_stream=%s  all stages %.3f ms  projection plan %.3f ms (min %.3f)  %.0f GB/s (survey bytes)  %.2f TFLOP/s FP64' % (
    os.environ.get('SINGLE_STREAM', '0'), np.mean(ta), np.mean(tp), np.min(tp), pp.algorithmic_bytes_survey / t / 1e9,
    pp.flops / t / 1e12))
