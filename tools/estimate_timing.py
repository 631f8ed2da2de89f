import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch, bench
from pylrbms_b200 import LRBMSReductor, discretize
a = bench.parse_args()
data, bases = bench.make_inputs(a)
d, _ = discretize(data); rd = LRBMSReductor(d, bases=bases).reduce()
n_mu = 10000
theta = torch.from_numpy(rd.thetas(bench.make_mus(a, 0, n_mu))).cuda()
u = torch.empty((n_mu, rd.n_red), dtype=torch.float64, device='cuda'); eta = torch.empty(n_mu, dtype=torch.float64, device='cuda'); info = torch.empty(n_mu, dtype=torch.int32, device='cuda')
rd.solve_device(theta, u, info)
for _ in range(3): rd.estimate_device(theta, u, eta)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rd.estimate_device(theta, u, eta); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print('LRBMS_EST_TMU', os.environ.get('LRBMS_EST_TMU'), 'estimate ms', np.mean(ts), 'eta sum', float(eta.sum()))
