"""Are two projections of the same bases bit-identical?  (same plan rerun; two reductors; different scratch placement)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pylrbms_b200 import LRBMSReductor, discretize
from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases
data = assemble_block_swipdg((8, 8), 32)
bases = make_local_bases(data, 20, seed=11)
bd = {'domain_%d' % i: bases[i] for i in range(64)}
d, _ = discretize(data)

def dense(rd):
    out = {}
    for name, op in rd.operators.items():
        for q, o in enumerate(getattr(op, 'operators', [op])):
            out[(name, q)] = o.to_dense()
    return out

r1 = LRBMSReductor(d, bases=bd); rd1 = r1.reduce(); A = dense(rd1)
r1.last_plan.run(); torch.cuda.synchronize(); A2 = dense(rd1)
r2 = LRBMSReductor(d, bases=bd); rd2 = r2.reduce(); B = dense(rd2)
junk = torch.empty(12345, dtype=torch.float64, device='cuda')       # shifts the allocations of the next reductor
r3 = LRBMSReductor(d, bases=bd); r3._proj_scratch = torch.full((400_000_000,), float('nan'), dtype=torch.float64, device='cuda')
rd3 = r3.reduce(); Cc = dense(rd3)
for tag, X in (('rerun', A2), ('second reductor', B), ('third reductor, NaN-poisoned scratch', Cc)):
    bad = [(k, int((A[k] != X[k]).sum()), float(np.nanmax(np.abs(A[k] - X[k])))) for k in A if not np.array_equal(A[k], X[k])]
    print(tag, ': differing operators', len(bad), bad[:6])
    if bad:
        k = bad[0][0]
        r, c = np.nonzero(A[k] != X[k])
        print('   rows', np.unique(r)[:20], 'cols', np.unique(c)[:20], 'n', len(r))
