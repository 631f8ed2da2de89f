"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): region exchange of disjoint reduced blocks and
the estimator-max gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pylrbms_b200.distributed import (exchange_regions, gather_estimator_max, mu_slice, owner_rank, rank_and_world,
                                              region_layout)
        assert rank_and_world() == (rank, world)
        # ---- offline: every rank fills the outputs of the subdomains it owns, then the regions are exchanged
        S = 7
        sizes = [3 * (s + 1) for s in range(S)]
        pending = [(owner_rank(s, S, world), sizes[s]) for s in range(S)]
        offsets, starts = region_layout(pending, world)
        buf = torch.zeros(int(starts[-1]), dtype=torch.float64)
        for s in range(S):
            if owner_rank(s, S, world) == rank:
                buf[offsets[s]:offsets[s] + sizes[s]] = float(s + 1)
        assert len({int(starts[r + 1] - starts[r]) for r in range(world)}) == 1      # equal strides: one all-gather
        exchange_regions(buf, starts)
        for s in range(S):
            assert torch.all(buf[offsets[s]:offsets[s] + sizes[s]] == float(s + 1))
        # foreign layout with unequal regions: the broadcast path
        starts2 = np.array([0, 5, 12])
        buf2 = torch.zeros(12, dtype=torch.float64)
        buf2[int(starts2[rank]):int(starts2[rank + 1])] = float(rank + 1)
        exchange_regions(buf2, starts2)
        assert torch.all(buf2[:5] == 1.0) and torch.all(buf2[5:] == 2.0)
        # a library handle on host tensors / a gloo group: the peer-memory path must decline, the collective carries it
        from pylrbms_b200.distributed import PeerStaging, exchange_kind
        buf3 = torch.zeros(int(starts[-1]), dtype=torch.float64)
        buf3[int(starts[rank]):int(starts[rank + 1])] = float(rank + 1)
        exchange_regions(buf3, starts, handle=object())
        for r in range(world):
            assert torch.all(buf3[int(starts[r]):int(starts[r + 1])] == float(r + 1))
        assert PeerStaging._current is None and exchange_kind().startswith('NCCL all_gather')
        # ---- online: eta over a global batch, sharded; max and arg-max must not depend on the sharding
        n_mu = 101
        eta_global = np.cos(np.arange(n_mu) * 0.37) ** 2
        eta_global[[17, 77]] = 2.5                                   # a tie across ranks: smallest index wins
        lo, hi = mu_slice(n_mu, rank, world)
        local = torch.from_numpy(eta_global[lo:hi])
        mx, am = local.max(), local.argmax()
        gmax, garg = gather_estimator_max(mx, am, lo)
        assert gmax == 2.5 and garg == 17
        np.save(os.path.join(result_dir, 'ok_%d.npy' % rank), np.array([gmax, garg]))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_and_estimator_max(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ('ok_%d.npy' % r)), [2.5, 17.0])


def test_single_process_fallthrough():
    """Without an initialised process group the helpers are no-ops / local."""
    from pylrbms_b200.distributed import exchange_regions, gather_estimator_max, rank_and_world
    assert rank_and_world() == (0, 1)
    buf = torch.arange(4.0)
    exchange_regions(buf, [0, 4])
    assert gather_estimator_max(torch.tensor(3.0), torch.tensor(5), 10) == (3.0, 15)
