"""C1 (OS2015, 4x4 subdomains, n_i = 1536, N = 8) through the whole hot path, small enough for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_c1.py
    compute-sanitizer --tool racecheck python tools/sanitize_c1.py

Runs the offline reduction, the shared-memory-window solve kernel, the band solver and the estimator on 24 parameters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pylrbms_b200 import LRBMSReductor, discretize
from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases

cells = int(os.environ.get('CELLS', '16'))
data = assemble_block_swipdg((4, 4), cells)
bases = make_local_bases(data, 8, seed=1001)
bd = {'domain_%d' % i: bases[i] for i in range(16)}
rd = LRBMSReductor(discretize(data)[0], bases=bd).reduce()
mus = np.linspace(0.1, 1.0, 24)
U, eta = rd.sweep(mus)
print('window kernel:', rd.solve_kernel_name, 'eta[0..2] =', eta[:3])
rd.set_solver('banded')
U2, eta2 = rd.sweep(mus)
print('band solver  :', rd.solve_kernel_name, 'max |du| =', float(np.abs(U.data - U2.data).max()))
