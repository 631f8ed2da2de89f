#!/usr/bin/env python
"""Benchmark of the LRBMS hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

Workload (BASELINE.json configs[1]): OS2015 academic multiscale example, 8x8
subdomains, 6144 fine DG dofs each, local basis size 20 (n_red = 1280), Q = 2 affine terms; online batch of
10 000 parameters per GPU, uniform in [0.1, 1].  Synthetic inputs: the operators come from the structured P1-SWIPDG
assembler in ``pylrbms_b200.swipdg_fixture`` (DUNE is not available), the bases are seeded and orthonormalised in
the local energy products.

A *step* = one sweep of the hot path over one parameter batch: assemble + factor + solve + estimate for every
parameter (``lrbms_online_solve`` + ``lrbms_online_estimate``) followed by the estimator max (``lrbms_eta_max``; for
N > 1 one NCCL all-reduce(max) -- the "estimator-max gather", the only collective of the path).  ``value`` counts
solves+estimates per second with the coefficient matrix already in HBM; ``e2e`` goes through the public API
(``ReducedModel.sweep_into``) from host parameters to host results.  The offline half of the metric (projection
HBM GB/s) is measured in the same run on the same workload and reported under ``offline``.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'online reduced solves+estimates/s (mu-batched)'
UNIT = 'solves+estimates/s'

# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` captures
# of the default workload (profiles/r01c_ncu_full_*_raw.csv); null for any other workload.
NCU_TRAFFIC_BYTES = {
    'solve_kernel_v2': 13.123e9 + 16.438e9,           # 10 000 parameters per launch
    'projection_plan': 3.995e9 + 0.826e9,             # SpMM stages >= 1 + the plan: 6 spmm + 4 project + 5 gram launches
}


# ----------------------------------------------------------------------------------------------------------
#  workload
# ----------------------------------------------------------------------------------------------------------

def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--subdomains', type=int, default=8, help='subdomains per direction (8 -> 8x8, config C2)')
    ap.add_argument('--cells', type=int, default=32, help='cells per subdomain and direction (32 -> n_i = 6144)')
    ap.add_argument('--basis', type=int, default=20, help='local basis size')
    ap.add_argument('--n-mu', type=int, default=10000, help='parameters per GPU and step')
    ap.add_argument('--cpu-sample', type=int, default=24, help='parameters per CPU-baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-offline', action='store_true')
    ap.add_argument('--seed', type=int, default=1002)
    ap.add_argument('--synthetic3d', default=None, metavar='HX,HY,HZ',
                    help='seeded synthetic operators with 3D structure (config C4 shape): --subdomains per direction, '
                         'HX x HY x HZ cells of 4 dofs per subdomain (16,16,12 -> n_i = 12288); offline measurements only')
    return ap.parse_args()


def workload_name(a):
    if a.synthetic3d:
        h = [int(x) for x in a.synthetic3d.split(',')]
        return 'synthetic 3D {0}x{0}x{0} subdomains, n_i={1}, N={2}, Q=2'.format(a.subdomains, 4 * h[0] * h[1] * h[2], a.basis)
    return 'OS2015 {0}x{0} subdomains, n_i={1}, N={2}, Q=2, {3} mu per GPU'.format(a.subdomains, 6 * a.cells ** 2, a.basis, a.n_mu)


def make_inputs(a):
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases
    if a.synthetic3d:
        from pylrbms_b200.synthetic_fixture import make_random_local_bases, synthetic_block_operators
        h = tuple(int(x) for x in a.synthetic3d.split(','))
        data = synthetic_block_operators((a.subdomains,) * 3, h, seed=a.seed)
        bases = make_random_local_bases(data, a.basis, seed=a.seed)
        return data, {'domain_%d' % i: bases[i] for i in range(data.num_subdomains)}
    data = assemble_block_swipdg((a.subdomains, a.subdomains), a.cells)
    bases = make_local_bases(data, a.basis, seed=a.seed)
    return data, {'domain_%d' % i: bases[i] for i in range(data.num_subdomains)}


def make_mus(a, rank, n):
    rng = np.random.default_rng(a.seed + 7919 * rank)
    return rng.uniform(0.1, 1.0, n)


def survey_flops_per_mu(sx, N, Q):
    """Algorithmic flops of one solve+estimate, SURVEY.md section 8d: assemble 2 Q B N^2, banded Cholesky n b^2,
    two triangular solves 4 n b, estimator sum_i 2 (d_i^2 + 2 (Q d_i)^2 + Q^2 N^2 + Q^2 N d_i + Q d_i)."""
    S = sx * sx
    n = S * N
    B = S + 2 * 2 * sx * (sx - 1)                 # diagonal + directed face-neighbour blocks
    b = (sx + 1) * N                              # scalar half bandwidth, lexicographic ordering
    est = 0
    for s in range(S):
        ix, iy = s % sx, s // sx
        nb = 1 + (ix > 0) + (ix < sx - 1) + (iy > 0) + (iy < sx - 1)
        d = nb * N
        est += 2 * (d * d + 2 * (Q * d) ** 2 + Q * Q * N * N + Q * Q * N * d + Q * d)
    solve = 2 * Q * B * N * N + n * b * b + 4 * n * b
    return dict(solve=float(solve), estimate=float(est), total=float(solve + est), n_red=n, blocks=B, half_bandwidth=b)


# ----------------------------------------------------------------------------------------------------------
#  clocks
# ----------------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
#  reference arm: the reference's CPU algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------------

def cpu_reference_model(a):
    """Offline phase with the oracle (NumPy/SciPy restatement of pylrbms on pyMOR); returns the reduced model."""
    from oracle import lrbms_oracle as O
    data, bases = make_inputs(a)
    t = time.perf_counter()
    d = O.build_discretization(data)
    red = O.LRBMSReductor(d, bases=bases)
    rd = red.reduce()
    return rd, time.perf_counter() - t


def cpu_online_step(rd, mus):
    """What the reference does per parameter: ``U = rd.solve(mu); eta = rd.estimate(U, mu)`` (online_enrichment.py:72-74)."""
    out = []
    for mu in mus:
        U = rd.solve(mu)
        out.append(rd.estimate(U, mu))
    return out


class VectorisedCpuModel:
    """SURVEY.md section 8d, CPU variant (ii): the same reduced model evaluated the way a careful CPU implementation
    would -- banded Cholesky (LAPACK dpbsv) on the block-banded reduced operator instead of a dense LU, and every estimator
    form restricted to the rows / columns it really touches instead of dense n_red-sized mat-vecs.  Reported next to the
    reference-faithful figure so that the GPU / CPU ratio is not inflated by the reference's dense storage."""

    def __init__(self, rd):
        import scipy.sparse as sp
        self.rd, est = rd, rd.estimator
        A = [np.asarray(o.matrix) for o in rd.operator.operators]
        n = A[0].shape[0]
        nz = np.nonzero(sum(np.abs(M) for M in A))
        self.b = int(np.max(np.abs(nz[0] - nz[1]))) if len(nz[0]) else 0
        self.ab = []
        for M in A:                                    # lower banded storage: ab[k, j] = M[j + k, j]
            ab = np.zeros((self.b + 1, n))
            for k in range(self.b + 1):
                ab[k, :n - k] = np.diagonal(M, -k)
            self.ab.append(ab)
        self.fr = [sp.csr_matrix(np.asarray(o.matrix)) for o in est.flux_reconstruction.operators]

        def restrict(M):
            M = np.asarray(M)
            r, c = np.nonzero(np.abs(M).sum(axis=1))[0], np.nonzero(np.abs(M).sum(axis=0))[0]
            return r, c, np.ascontiguousarray(M[np.ix_(r, c)])
        self.sub = []
        for ii, s_ in enumerate(est.subdomains):
            ops = rd.operators
            self.sub.append(dict(
                nc=restrict(ops['nc_%d' % s_].matrix), r_fd=restrict(ops['r_fd_%d' % s_]._array.data),
                r_dd=restrict(ops['r_dd_%d' % s_].matrix), df_bb=restrict(ops['df_bb_%d' % s_].matrix),
                df_aa=[(c, restrict(o.matrix)) for o, c in zip(ops['df_aa_%d' % s_].operators, ops['df_aa_%d' % s_].coefficients)],
                df_ab=[(c, restrict(o.matrix)) for o, c in zip(ops['df_ab_%d' % s_].operators, ops['df_ab_%d' % s_].coefficients)]))

    @staticmethod
    def _form(x, rcm, y):
        r, c, M = rcm
        return x[r] @ (M @ y[c])

    def step(self, mus):
        from scipy.linalg import solveh_banded
        rd, est = self.rd, self.rd.estimator
        out = []
        for mu_ in mus:
            mu = rd.parse_parameter(mu_)
            th = [c.evaluate(mu) for c in rd.operator.coefficients]
            ab = th[0] * self.ab[0]
            for q in range(1, len(th)):
                ab = ab + th[q] * self.ab[q]
            f = rd.rhs.as_source_array(mu).data[0]
            u = solveh_banded(ab, f, lower=True)
            thf = [c.evaluate(mu) for c in est.flux_reconstruction.coefficients]
            ur = sum(t * (S @ u) for t, S in zip(thf, self.fr))
            S_ = len(self.sub)
            nc, r, df = np.zeros(S_), np.zeros(S_), np.zeros(S_)
            for ii, T in enumerate(self.sub):
                nc[ii] = self._form(u, T['nc'], u)
                rr, cc, M = T['r_fd']
                r[ii] = est.local_eta_rf_squared[ii] - 2.0 * float(M[0] @ ur[cc]) + self._form(ur, T['r_dd'], ur)
                df[ii] = sum(c.evaluate(mu) * self._form(u, R, u) for c, R in T['df_aa']) + self._form(ur, T['df_bb'], ur) + \
                    2.0 * sum(c.evaluate(mu) * self._form(u, R, ur) for c, R in T['df_ab'])
                r[ii] *= (1.0 / np.pi ** 2) / est.min_diffusion_evs[ii] * est.subdomain_diameters[ii] ** 2
            a_bar = est.alpha(est.lambda_coeffs, mu, est.mu_bar)
            g_bar = est.gamma(est.lambda_coeffs, mu, est.mu_bar)
            a_hat = est.alpha(est.lambda_coeffs, mu, est.mu_hat)
            out.append((np.sqrt(g_bar) * np.linalg.norm(nc) + np.linalg.norm(r + df) / np.sqrt(a_hat)) / np.sqrt(a_bar))
        return out


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core the BLAS can take."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get('num_threads', 1) for p in threadpool_info()] + [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    use_all_host_threads()
    rd, t_off = cpu_reference_model(a)
    mus = make_mus(a, 0, a.cpu_sample)
    for _ in range(max(1, a.warmup) if a.warmup else 0):
        cpu_online_step(rd, mus[:max(1, a.cpu_sample // 4)])
    t = time.perf_counter()
    for _ in range(a.steps):
        cpu_online_step(rd, mus)
    dt = time.perf_counter() - t
    value = a.cpu_sample * a.steps / dt
    cores = cpu_threads()
    sample = '{} parameters per step of the same reduced model (n_red={}), dense unblocked operators, numpy.linalg.solve + ' \
             '6 quadratic forms per subdomain per parameter'.format(a.cpu_sample, rd.solution_space.dim)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * dt / a.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(a), 'cpu_sample_mu_per_step': a.cpu_sample,
                   'note': 'pyMOR/DUNE are not installable here; this is the NumPy/SciPy restatement of the reference path (oracle/)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                         'offline_reduce_s': t_off},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
#  B200 arm
# ----------------------------------------------------------------------------------------------------------

def measure_fp64_gemm_peak(torch):
    """cuBLAS DGEMM 4096^3, best of 5 (burst) -- the FP64 tensor-pipe denominator; MEASURED_PEAKS.json has no FP64 figure."""
    n = 4096
    a = torch.randn((n, n), dtype=torch.float64, device='cuda')
    b = torch.randn((n, n), dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (measured)'
    except Exception:
        return 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'


def run_b200(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from pylrbms_b200 import build
    build.build()
    from pylrbms_b200 import LRBMSReductor, discretize

    # ---- setup (untimed): host assembly, upload, offline reduction -- every rank holds the full (small) reduced model
    data, bases = make_inputs(a)
    d, _ = discretize(data)
    reductor = LRBMSReductor(d, bases=bases)
    t0 = time.perf_counter()
    rd = reductor.reduce()
    torch.cuda.synchronize()
    t_reduce_first = time.perf_counter() - t0
    planner = reductor.last_plan
    if not a.synthetic3d:
        rd.online_plan                                 # build the online plan (symbolic phase + tile upload)

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device='cuda')     # 256 MB > 126 MB L2
    fp64_peak = measure_fp64_gemm_peak(torch)

    def flush_l2():
        flush.fill_(1.0)

    # ---- offline half: the batched projection of every operator (kernel-only, inputs resident in HBM)
    offline = None
    if not a.no_offline:
        st = planner.stats()
        for _ in range(max(3, a.warmup)):
            planner.run()
        times_all, times_proj = [], []
        for _ in range(max(3, a.steps)):
            flush_l2()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            planner.spmm_plans[0].run()        # stage 0: Oswald / flux-reconstruction images of the bases (rows a2, a3)
            e1.record()
            for p in planner.spmm_plans[1:]:   # later stages belong to the projection: D R, and A^T L of the narrow-left chains
                p.run()
            planner.project_plan.run()
            e2.record()
            e2.synchronize()
            times_all.append(e0.elapsed_time(e2)); times_proj.append(e1.elapsed_time(e2))
        hbm_peak, hbm_src = measured_peaks()
        pp = planner.project_plan
        t_proj = float(np.mean(times_proj)) * 1e-3
        # Algorithmic work = SURVEY.md section 8d formula on the operators as the reference defines them (every chain applied
        # matrix by matrix to the right-hand array).  The plan that runs does less: products of chained sparse matrices are
        # formed once on the host (r_dd: Gram over the m_i flux dofs instead of the n_i DG dofs) and narrow-left chains apply
        # the matrix to the narrow side; its own count is reported as *_executed.
        acct = LRBMSReductor(d, bases=bases)
        acct.fuse_chains, acct.narrow_left = False, False
        acct_plan = acct.build_plan()
        alg_bytes = acct_plan.project_plan.algorithmic_bytes_survey
        alg_flops = acct_plan.project_plan.flops
        alg_descs = acct_plan.n_project_descs
        del acct, acct_plan
        gbs = alg_bytes / t_proj / 1e9
        tfs = alg_flops / t_proj / 1e12
        ai = alg_flops / alg_bytes
        tensor_bound = ai > fp64_peak * 1e12 / (hbm_peak * 1e9)      # arithmetic intensity above the HBM / FP64 crossover
        roof = {'bound': 'tensor' if tensor_bound else 'hbm', 'kernel': 'projection plan (all buckets of one run: spmm_kernel, '
                                                                        'project_kernel, gram_kernel)',
                'achieved': tfs if tensor_bound else gbs, 'peak': fp64_peak if tensor_bound else hbm_peak,
                'unit': 'TFLOP/s' if tensor_bound else 'GB/s',
                'frac': tfs / fp64_peak if tensor_bound else gbs / hbm_peak,
                'traffic': NCU_TRAFFIC_BYTES['projection_plan'] if (a.subdomains, a.cells, a.basis, a.synthetic3d) == (8, 32, 20, None) else None,
                'peak_source': ('cuBLAS DGEMM 4096^3 measured in this run (FP64 tensor pipe)' if tensor_bound else hbm_src),
                'arithmetic_intensity_flop_per_byte': ai,
                'algorithmic_bytes_survey_formula': alg_bytes, 'flops': alg_flops,
                'bytes_executed_plan': pp.algorithmic_bytes_survey, 'flops_executed_plan': pp.flops,
                'fp64_tflops_executed': pp.flops / t_proj / 1e12, 'hbm_gbs': gbs, 'hbm_peak_gbs': hbm_peak, 'hbm_frac': gbs / hbm_peak, 'hbm_peak_source': hbm_src,
                'fp64_tflops': tfs, 'fp64_peak_tflops': fp64_peak, 'fp64_frac': tfs / fp64_peak,
                'note': 'bound = whichever of the two rooflines binds at this arithmetic intensity (crossover %.1f flop/B); both '
                        'fractions are reported' % (fp64_peak * 1e12 / (hbm_peak * 1e9))}
        offline = {
            'metric': 'offline projection HBM GB/s', 'unit': 'GB/s', 'value': gbs,
            'ms_all_stages': float(np.mean(times_all)), 'ms_projection': float(np.mean(times_proj)),
            'ms_projection_covers': 'SpMM stages >= 1 (divergence images, A^T L of narrow-left chains) + the projection plan; '
                                    'stage 0 (Oswald / flux-reconstruction images of the bases) is the rest of ms_all_stages',
            'projection_descriptors': planner.n_project_descs, 'spmm_descriptors': planner.n_spmm_descs,
            'launches_per_reduce': st['launches'], 'roofline': roof,
            'first_reduce_incl_planning_s': t_reduce_first,
        }
    # ---- N > 1: the subdomain-sharded offline half (strong scaling: the same operator set split over the ranks, each rank
    #      projects the blocks of its strip of subdomains; the reduced regions are exchanged over NCCL afterwards)
    if world > 1 and offline is not None:
        red_s = LRBMSReductor(d, bases=bases, shard=True)
        red_s.reduce()
        ps = red_s.last_plan
        for _ in range(3):
            ps.run(); ps.exchange()
        t_run, t_all = [], []
        for _ in range(max(3, a.steps)):
            flush_l2()
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            ps.run()
            e1.record()
            ps.exchange()
            e2.record()
            e2.synchronize()
            t_run.append(e0.elapsed_time(e1)); t_all.append(e0.elapsed_time(e2))
        tt = torch.tensor([float(np.mean(t_run)), float(np.mean(t_all))], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        offline['sharded'] = {
            'n_gpus': world, 'scaling': 'strong', 'partition': 'contiguous strips of subdomains per rank',
            'ms_all_stages_max_over_ranks': float(tt[0].item()), 'ms_with_exchange_max_over_ranks': float(tt[1].item()),
            'ms_all_stages_1gpu_unsharded_this_rank': offline['ms_all_stages'],
            'speedup_vs_unsharded': offline['ms_all_stages'] / float(tt[0].item()),
            'exchange': 'one NCCL broadcast per rank region (all-gather of disjoint reduced blocks), %d doubles in total' % int(ps.out.numel()),
            'projection_descriptors_this_rank': ps.n_project_descs,
        }
    if a.synthetic3d:
        # auxiliary measurement (config C4 shape): the offline half only; the driver's bench line is the default workload
        if rank == 0:
            print(json.dumps({'metric': 'offline projection HBM GB/s', 'value': offline['value'], 'unit': 'GB/s', 'n_gpus': world,
                              'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': workload_name(a)},
                              'higher_is_better': True, 'offline': offline}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- online half
    n_mu = a.n_mu
    mus = make_mus(a, rank, n_mu)
    theta = torch.from_numpy(rd.thetas(mus)).cuda()
    S = len(rd.block_dims)
    u = torch.empty((n_mu, rd.n_red), dtype=torch.float64, device='cuda')
    eta = torch.empty(n_mu, dtype=torch.float64, device='cuda')
    info = torch.empty(n_mu, dtype=torch.int32, device='cuda')

    def step():
        rd.solve_device(theta, u, info)
        ev_mid = torch.cuda.Event(enable_timing=True); ev_mid.record()
        rd.estimate_device(theta, u, eta)
        mx, am = rd.eta_max_device(eta)
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        return ev_mid, mx

    for _ in range(max(3, a.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_step, t_solve = [], []
    barrier()
    for _ in range(a.steps):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ev_mid, mx = step()
        e1.record()
        e1.synchronize()
        t_step.append(e0.elapsed_time(e1)); t_solve.append(e0.elapsed_time(ev_mid))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if int(info.max().item()) != 0:
        raise SystemExit('bench.py: a reduced system was flagged as not positive definite')
    total_ms = torch.tensor([float(np.sum(t_step))], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = world * n_mu * a.steps / total_s

    # ---- end to end through the public API: host parameters -> host (U, eta), copies inside the timed region
    u_host = torch.empty((n_mu, rd.n_red), dtype=torch.float64).pin_memory()
    eta_host = torch.empty(n_mu, dtype=torch.float64).pin_memory()
    for _ in range(2):
        rd.sweep_into(mus, u_host, eta_host)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(a.steps, 5))
    for _ in range(e2e_steps):
        rd.sweep_into(mus, u_host, eta_host)
        if world > 1:
            m = torch.tensor([float(eta_host.max())], dtype=torch.float64, device='cuda')
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n_mu * e2e_steps / float(e2e_s.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (solve_kernel: FP64 tensor-core tile Cholesky)
    fl = survey_flops_per_mu(a.subdomains, a.basis, data.Q)
    solve_s = float(np.mean(t_solve)) * 1e-3
    achieved = fl['solve'] * n_mu / solve_s / 1e12
    from pylrbms_b200._lib import Symbolic
    default_workload = (a.subdomains, a.cells, a.basis, a.n_mu) == (8, 32, 20, 10000)
    solve_name = 'solve_kernel_v2' if os.environ.get('LRBMS_SOLVE_V1', '0') in ('', '0') else 'solve_kernel'
    roofline = {'bound': 'tensor', 'kernel': solve_name, 'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                'frac': achieved / fp64_peak,
                'traffic': NCU_TRAFFIC_BYTES.get(solve_name) if default_workload else None,
                'peak_source': 'cuBLAS DGEMM 4096^3 measured in this run (FP64; MEASURED_PEAKS.json has no FP64 figure)',
                'algorithmic_flops_per_mu': fl['solve'], 'ms_per_launch': 1e3 * solve_s,
                'share_of_step': float(np.sum(t_solve) / np.sum(t_step))}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(3, a.warmup),
        'ms_per_step': 1e3 * total_s / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(a), 'n_red': fl['n_red'], 'reduced_blocks': fl['blocks'],
                   'l2': 'flushed between timed steps (256 MB write); factor scratch per step also exceeds L2',
                   'parallelism': 'mu-sharded x{}'.format(world), 'flops_per_mu_survey': fl['total']},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(theta.numel() * 8),
                'd2h_bytes_per_step': int((u.numel() + eta.numel()) * 8), 'api': 'ReducedModel.sweep_into(mus, u_host, eta_host)'},
        'gpu_launches': int(a.steps * 4),
        'roofline': roofline,
        'offline': offline,
    }
    if not a.no_cpu_baseline and world == 1:
        line['cpu_baseline'] = cpu_baseline(a)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a):
    use_all_host_threads()
    rd, t_off = cpu_reference_model(a)
    mus = make_mus(a, 0, a.cpu_sample)
    cpu_online_step(rd, mus[:2])
    t = time.perf_counter()
    cpu_online_step(rd, mus)
    dt = time.perf_counter() - t
    # variant (ii) of SURVEY.md section 8d: banded Cholesky + restricted estimator forms; checked against variant (i)
    vm = VectorisedCpuModel(rd)
    ref = cpu_online_step(rd, mus[:2])
    got = vm.step(mus[:2])
    assert np.allclose(got, ref, rtol=1e-9, atol=0.0), (got, ref)
    n_vec = 8 * a.cpu_sample
    mus_v = make_mus(a, 1, n_vec)
    t = time.perf_counter()
    vm.step(mus_v)
    dt_v = time.perf_counter() - t
    return {'value': a.cpu_sample / dt, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
            'sample': '{} parameters of the same workload, one at a time as the reference does (dense unblocked operators, '
                      'numpy.linalg.solve, 6 quadratic forms per subdomain)'.format(a.cpu_sample),
            'offline_reduce_s': t_off,
            'vectorised': {'value': n_vec / dt_v, 'unit': UNIT, 'sample': '{} parameters'.format(n_vec),
                           'what': 'same reduced model, banded Cholesky (scipy.linalg.solveh_banded, half-bandwidth {}) and estimator '
                                   'forms restricted to the rows / columns they touch; agrees with the reference-faithful variant to '
                                   '1e-9'.format(vm.b)}}


if __name__ == '__main__':
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)
