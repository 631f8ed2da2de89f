"""Where does the eta discrepancy at C2 come from?  (u vs estimator forms; cancellation of df / r)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pylrbms_b200 import LRBMSReductor, discretize
from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases
from oracle import lrbms_oracle as O
from oracle.parity import compare_online, reference_online
from oracle.pymor_like import VA

data = assemble_block_swipdg((8, 8), 32)
bases = make_local_bases(data, 20, seed=1002)
bd = {'domain_%d' % i: bases[i] for i in range(64)}
rd = LRBMSReductor(discretize(data)[0], bases=bd).reduce()
rd_ref = O.LRBMSReductor(O.build_discretization(data), bases=bd).reduce()
mus = np.array([0.1, 0.55, 1.0])
ref = reference_online(rd_ref, mus)
print('errors vs oracle:', compare_online(rd, rd_ref, mus, ref=ref))
U, eta, parts, ind = rd.sweep(mus, decompose=True)
U_ref, eta_ref, parts_ref, ind_ref, As = ref
for k, mu in enumerate(mus):
    # oracle estimator on the GPU solution: separates the solve from the estimator kernel
    Ug = rd_ref.solve(mu); Ug.data[0][:] = U.data[k]
    e_g, p_g, _ = rd_ref.estimate(Ug, mu, decompose=True)
    print('mu', mu, 'eta_gpu', eta[k], 'eta_ref', eta_ref[k], 'oracle-estimator(U_gpu)', e_g,
          '\n   rel(eta_gpu, eta_ref) %.2e  rel(oracle(U_gpu), eta_ref) %.2e  rel(eta_gpu, oracle(U_gpu)) %.2e' % (
              abs(eta[k] - eta_ref[k]) / eta_ref[k], abs(e_g - eta_ref[k]) / eta_ref[k], abs(eta[k] - e_g) / e_g))
    for name, j in (('nc', 0), ('r', 1), ('df', 2)):
        print('   %-3s max|.| %.3e  max diff gpu-ref %.3e  sum %.6e' % (name, np.abs(parts_ref[j][:, k]).max(),
              np.abs(parts[j][:, k] - parts_ref[j][:, k]).max(), parts_ref[j][:, k].sum()))
    e = U.data[k] - U_ref[k]
    print('   u energy rel %.2e  max rel %.2e  cond(A) %.3e' % (np.sqrt(e @ As[k] @ e) / np.sqrt(U_ref[k] @ As[k] @ U_ref[k]),
          np.abs(e).max() / np.abs(U_ref[k]).max(), np.linalg.cond(As[k])))
est = rd.estimator
print('||f||^2 * scale: max %.3e' % np.abs(np.asarray(est.local_eta_rf_squared) * np.asarray(est.r_scale())).max())
# magnitude of the df terms for subdomain 27: u^T AA u etc.
ops = rd_ref.operators
u = U_ref[0]
mu_p = rd_ref.parse_parameter(mus[0])
aa = ops['df_aa_27'].assemble(mu_p).matrix
print('df_aa term (sub 27):', float(u @ aa @ u), ' df_27 =', parts_ref[2][27, 0])
