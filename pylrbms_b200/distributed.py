"""Multi-GPU plumbing of the LRBMS hot path: one process per GPU, ``torch.distributed`` (NCCL on the B200 box, gloo
in the CPU tests).  The path shards without any data-path collective:

* offline -- contiguous strips of subdomains per rank (the reference's own decomposition, ``subdomains_on_rank``,
  ``estimators.py:40,70``); the reduced blocks of every rank are disjoint, so the exchange is an all-gather
  (one in-place ``all_gather_into_tensor`` over equal-stride rank regions), replacing the dead ``Allreduce(SUM)`` of zero-padded blocks at ``reductor.py:93``;
* online -- the parameter batch is split evenly; the only exchange per sweep is the estimator maximum and its
  arg-max (replacing ``mpi_norm``'s sum-reduce, ``estimators.py:100-101``, which disappears because every rank holds
  all subdomains of the small reduced model).

Everything here works on tensors of whatever device the process group's backend wants.
"""
from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed():
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_and_world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def owner_rank(owner, num_subdomains, world):
    """Contiguous strips: subdomain ``owner`` belongs to rank ``floor(owner * world / S)``."""
    return min(world - 1, int(owner) * world // max(1, int(num_subdomains)))


def subdomains_on_rank(num_subdomains, rank, world):
    return [s for s in range(num_subdomains) if owner_rank(s, num_subdomains, world) == rank]


def region_layout(pending, world):
    """``pending``: list of ``(owner_rank, size)`` output requests in planning order.  Returns ``(offsets, starts)``:
    the offset of every request inside one buffer in which each rank's outputs form one contiguous region, and
    the ``world + 1`` region boundaries.  With more than one rank all regions get the same length (the largest, rounded up
    to 32 doubles) so that the exchange is ONE in-place all-gather instead of one broadcast per rank."""
    region = np.zeros(world, dtype=np.int64)
    for r, size in pending:
        region[r] += size
    if world > 1:
        region[:] = (int(region.max()) + 31) // 32 * 32
    starts = np.concatenate([[0], np.cumsum(region)]).astype(np.int64)
    cursor = starts[:-1].copy()
    offsets = np.zeros(len(pending), dtype=np.int64)
    for t, (r, size) in enumerate(pending):
        offsets[t] = cursor[r]
        cursor[r] += size
    return offsets, starts


def exchange_regions(buffer, starts):
    """All-gather of disjoint contiguous regions of ``buffer`` (rank ``r`` owns ``buffer[starts[r]:starts[r+1]]``).

    Equal-length regions (what ``region_layout`` produces): one in-place ``all_gather_into_tensor`` -- a single NCCL
    collective over NVLink / NVSwitch, launch-latency-bound at these sizes (tens of MB), where eight serialised
    broadcasts cost eight launches.  Unequal regions (foreign layouts) fall back to one broadcast per rank."""
    if not is_distributed():
        return
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [int(starts[r + 1]) - int(starts[r]) for r in range(world)]
    if len(set(sizes)) == 1 and sizes[0] > 0 and int(starts[world]) == buffer.numel():
        mine = buffer[int(starts[rank]):int(starts[rank + 1])]
        try:
            dist.all_gather_into_tensor(buffer, mine)
            return
        except (RuntimeError, NotImplementedError):     # a backend without the flat all-gather (older gloo)
            parts = [buffer[int(starts[r]):int(starts[r + 1])] for r in range(world)]
            dist.all_gather(parts, mine.clone())
            return
    works = []
    for r in range(world):
        a, b = int(starts[r]), int(starts[r + 1])
        if b > a:
            works.append(dist.broadcast(buffer[a:b], src=r, async_op=True))
    for w in works:
        w.wait()


def mu_slice(n_mu, rank, world):
    """Even split of a parameter batch: rank ``r`` handles ``[lo, hi)``."""
    lo = (n_mu * rank) // world
    hi = (n_mu * (rank + 1)) // world
    return lo, hi


def gather_estimator_max(local_max, local_argmax, mu_offset):
    """The "estimator-max gather": global maximum of eta over all ranks and the global index of the maximiser.

    ``local_max`` (1 double) and ``local_argmax`` (1 int64, index inside this rank's slice) live on the backend's
    device; ``mu_offset`` is this rank's slice start.  One all-gather of ``world`` (max, index) pairs per sweep; ties
    go to the smallest global index so the result does not depend on the number of ranks."""
    import torch
    gidx = local_argmax.to(torch.float64) + float(mu_offset)
    pair = torch.stack([local_max.reshape(()).to(torch.float64), gidx.reshape(())])
    if not is_distributed():
        return float(pair[0].item()), int(pair[1].item())
    dist = _dist()
    world = dist.get_world_size()
    out = torch.empty((world, 2), dtype=torch.float64, device=pair.device)
    dist.all_gather_into_tensor(out, pair.reshape(1, 2))
    out = out.cpu().numpy()
    best = max(range(world), key=lambda r: (out[r, 0], -out[r, 1]))
    return float(out[best, 0]), int(out[best, 1])
