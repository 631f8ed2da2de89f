"""Multi-GPU check (run under torchrun, one rank per GPU): subdomain-sharded offline reduction + region exchange and
the mu-sharded online sweep with the estimator-max gather must reproduce the single-GPU results bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from pylrbms_b200 import LRBMSReductor, discretize
    from pylrbms_b200.swipdg_fixture import assemble_block_swipdg, make_local_bases
    sx, cells = int(os.environ.get('SUBDOMAINS', '4')), int(os.environ.get('CELLS', '8'))
    data = assemble_block_swipdg((sx, sx), cells)
    nb = os.environ.get('BASIS')
    bases = make_local_bases(data, int(nb) if nb else [5 + (i % 4) for i in range(sx * sx)], seed=11)
    bd = {'domain_%d' % i: bases[i] for i in range(sx * sx)}
    d, _ = discretize(data)
    red_u = LRBMSReductor(d, bases=bd)
    rd_u = red_u.reduce()
    red_s = LRBMSReductor(d, bases=bd, shard=True)
    rd_s = red_s.reduce()
    torch.cuda.synchronize()
    # the sharded planner groups outputs per owner rank, so compare operator by operator
    worst = 0.0
    for name in rd_u.operators:
        a, b = rd_u.operators[name], rd_s.operators[name]
        ta = a.operators if hasattr(a, 'operators') else [a]
        tb = b.operators if hasattr(b, 'operators') else [b]
        for x, y in zip(ta, tb):
            A, B = x.to_dense(), y.to_dense()
            assert A.shape == B.shape
            diff = float(np.abs(A - B).max())
            if diff > 0 and worst == 0.0:
                print('first difference:', name, diff, 'scale', float(np.abs(A).max()), 'nonzero entries differing', int((A != B).sum()),
                      'rows_per_cta hint', red_u.last_plan.unit_rows, red_s.last_plan.unit_rows)
            worst = max(worst, diff)
    assert worst == 0.0, worst
    mine = red_s.last_plan.n_project_descs
    total = red_u.last_plan.n_project_descs
    cnt = torch.tensor([mine], device='cuda')
    dist.all_reduce(cnt)
    assert int(cnt.item()) == total, (int(cnt.item()), total)
    # online: sharded sweep vs full sweep
    mus = np.random.default_rng(3).uniform(0.1, 1.0, 1001)
    U, eta = rd_u.sweep(mus)
    Ul, etal, (lo, hi), eta_max, argmax = rd_s.sweep_sharded(mus)
    assert np.array_equal(etal.cpu().numpy(), eta[lo:hi])
    assert np.array_equal(Ul.data, U.data[lo:hi])
    assert eta_max == eta.max() and argmax == int(np.argmax(eta))
    if rank == 0:
        print('sharded check ok: world', world, 'descriptors on rank 0:', mine, 'of', total, 'eta_max', eta_max, 'argmax', argmax)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
