"""Multi-GPU check of the peer-memory exchange (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_peer_exchange.py

Every iteration fills the ranks' regions with fresh pseudo-random doubles (so stale staging data cannot pass), exchanges
them through ``PeerStaging`` and compares the gathered buffer bit for bit with an NCCL all-gather of the same regions;
sizes grow on the way (the staging buffer is re-allocated collectively).  Then both paths are timed with CUDA events
(max over ranks).  Exit code 0 = identical everywhere."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
    from pylrbms_b200._lib import Handle
    from pylrbms_b200.distributed import PeerStaging, exchange_kind, exchange_regions, region_layout
    h = Handle.get()
    ok = True
    report = {'world': world, 'cases': []}
    for case, per_rank in enumerate([1000, 70_000, 700_000, 300_000, 2_800_000]):
        pending = [(r, per_rank + 17 * r) for r in range(world)]
        _, starts = region_layout(pending, world)
        total = int(starts[-1])
        for it in range(4):
            g = torch.Generator(device='cuda').manual_seed(1000 * case + 10 * it + rank)
            buf = torch.zeros(total, dtype=torch.float64, device='cuda')
            a, b = int(starts[rank]), int(starts[rank + 1])
            buf[a:b] = torch.rand(b - a, generator=g, dtype=torch.float64, device='cuda')
            ref = buf.clone()
            exchange_regions(ref, starts)                 # NCCL
            exchange_regions(buf, starts, handle=h)       # peer memory (or NCCL again if unavailable)
            torch.cuda.synchronize()
            same = bool(torch.equal(buf, ref)) and bool((ref != 0).any())
            ok = ok and same
        # timing
        buf = torch.rand(total, dtype=torch.float64, device='cuda')
        times = {}
        for name, hh in (('peer', h), ('nccl', None)):
            ts = []
            for it in range(13):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                exchange_regions(buf, starts, handle=hh)
                e1.record()
                e1.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1))
            t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times[name] = float(t.item())
        report['cases'].append({'doubles_per_rank': int(starts[1]), 'ms_peer': times['peer'], 'ms_nccl': times['nccl']})
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report['identical'] = bool(flag.item() == 1)
    report['exchange'] = exchange_kind()
    report['peer_disabled_reason'] = PeerStaging._disabled
    if rank == 0:
        print(json.dumps(report))
    dist.destroy_process_group()
    sys.exit(0 if report['identical'] else 1)


if __name__ == '__main__':
    main()
