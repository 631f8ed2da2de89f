import os, sys, numpy as np, ctypes as C
os.environ['LRBMS_SOLVE_TIMING']='1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from pylrbms_b200 import build; build.build()
from pylrbms_b200 import LRBMSReductor, discretize
a=bench.parse_args()
data,bases=bench.make_inputs(a)
d,_=discretize(data); rd=LRBMSReductor(d,bases=bases).reduce()
n_mu=148*8
theta=torch.from_numpy(rd.thetas(bench.make_mus(a,0,n_mu))).cuda()
for _ in range(2): rd.solve_device(theta)
torch.cuda.synchronize()
plan=rd.online_plan
out=np.zeros(128,dtype=np.int64)
n=plan.handle.lib.lrbms_online_debug_timing(plan.p, out.ctypes.data, 128)
t=out.reshape(16,8)/8.0   # per mu (8 mu per CTA)
names=['meta','chain|Y','ubar','X','endbar','epi','back','store']
print('cycles per mu, per warp:')
print('warp '+' '.join('%9s'%n for n in names)+'   total')
for w in range(16): print('%4d '%w+' '.join('%9.0f'%v for v in t[w])+'  %9.0f'%t[w].sum())
print('per column: ', ' '.join('%s=%.0f'%(n, t[1:,k].mean()/160) for k,n in enumerate(names[:6])), ' warp0 chain=%.0f'%(t[0,1]/160))
