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


class PeerStaging:
    """Symmetric (peer-mapped) staging buffer of the offline exchange on one NVLink / NVSwitch node.

    Every rank allocates the same ``2 * capacity`` doubles with ``torch.distributed._symmetric_memory`` (CUDA VMM
    allocations whose handles are exchanged once, at the rendezvous); afterwards each rank holds device pointers to
    every peer's buffer and, where the fabric offers it, ONE multicast address that the NVSwitch replicates to all
    GPUs.  An exchange is then: ``peer_push_kernel`` (this rank's region -> the same offset of every peer's staging
    half, plain stores over NVLink or a single multicast store stream), a device-side barrier over the signal pads,
    and a local copy of the other ranks' regions out of the staging half -- no NCCL launch on the path.

    The two halves alternate: a rank can only push into half ``k % 2`` for exchange ``k + 2`` after it passed the
    barrier of exchange ``k + 1``, which every peer reaches (in stream order) only after its copy-out of exchange
    ``k`` -- so a fast rank never overwrites staging data a slow rank still reads.

    Construction and growth are collective: all ranks must request the same capacity in the same order (they do:
    the region layout is a global function of the plan)."""

    _current = None          # one staging buffer per process, grown on demand
    _disabled = None         # reason string once the peer path has been found unusable

    def __init__(self, numel, mode):
        import torch
        import torch.distributed._symmetric_memory as symm
        dist = _dist()
        self.capacity = (int(numel) + 31) // 32 * 32
        dev = torch.device('cuda', torch.cuda.current_device())
        self.buf = symm.empty(2 * self.capacity, dtype=torch.float64, device=dev)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        mc = 0
        # two ranks: one destination either way, and plain peer stores measured faster than the multicast path
        if mode == 'multicast' or (mode != 'unicast' and self.world > 2):
            try:
                mc = int(self.hdl.multicast_ptr) if self.hdl.has_multicast_support else 0
            except (RuntimeError, AttributeError):
                mc = 0
        if mode == 'multicast' and not mc:
            raise RuntimeError('LRBMS_PEER_EXCHANGE=multicast: the symmetric allocation has no multicast address')
        self.multicast_ptr = mc
        self.exchanges = 0

    @property
    def kind(self):
        return 'nvswitch multicast stores' if self.multicast_ptr else 'nvlink peer stores'

    @classmethod
    def acquire(cls, numel):
        """The process-wide staging buffer with room for ``numel`` doubles per half, or ``None`` when the peer path
        is switched off (``LRBMS_PEER_EXCHANGE=0``), the backend is not NCCL, or symmetric memory cannot be set up
        (the ranks agree on that with one all-reduce, so no rank is left waiting in a rendezvous)."""
        import os
        mode = os.environ.get('LRBMS_PEER_EXCHANGE', 'auto').lower()
        if mode in ('0', 'off', 'no', 'false'):
            return None
        if cls._disabled is not None:
            return None
        if cls._current is not None and cls._current.capacity >= numel:
            return cls._current
        import torch
        dist = _dist()
        if dist.get_backend() != 'nccl' or not torch.cuda.is_available():
            cls._disabled = 'process group backend is not nccl'
            return None
        new, why = None, ''
        try:
            grow = numel if cls._current is None else max(numel, 2 * cls._current.capacity)
            new = cls(grow, mode)
        except Exception as e:      # noqa: BLE001 -- any failure means "no peer path"; NCCL carries the exchange
            why = f'{type(e).__name__}: {e}'
        ok = torch.tensor([1 if new is not None else 0], dtype=torch.int32, device='cuda')
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            cls._disabled = why or 'symmetric memory unavailable on another rank'
            cls._current = None
            return None
        cls._current = new
        return new

    def exchange(self, handle, buffer, starts):
        """``buffer[starts[r]:starts[r+1]]`` of rank ``r`` -> the same slice of ``buffer`` on every rank."""
        half = self._push(handle, buffer, starts)
        self.hdl.barrier(channel=0)
        self._collect(buffer, starts, half)

    def _push(self, handle, buffer, starts):
        """Store this rank's region into the current staging half of every peer (or the multicast address); returns the half."""
        import ctypes as C
        from . import _lib
        rank, world = self.rank, self.world
        a, b = int(starts[rank]), int(starts[rank + 1])
        half = self.exchanges & 1
        self.exchanges += 1
        base = 8 * (half * self.capacity + a)
        if self.multicast_ptr:
            dsts = [self.multicast_ptr + base]
        else:
            dsts = [self.peer_ptrs[p] + base for p in range(world) if p != rank]
        arr = (C.c_uint64 * len(dsts))(*dsts)
        handle.check(handle.lib.lrbms_peer_push(handle.h, C.c_void_p(buffer.data_ptr() + 8 * a), 8 * (b - a), len(dsts),
                                                C.cast(arr, C.c_void_p), 1 if self.multicast_ptr else 0,
                                                _lib.current_stream_ptr()))
        return half

    def _collect(self, buffer, starts, half):
        """After the barrier: the other ranks' regions out of this rank's staging half."""
        total = int(starts[self.world])
        a, b = int(starts[self.rank]), int(starts[self.rank + 1])
        stage = self.buf[half * self.capacity: half * self.capacity + total]
        if a > 0:
            buffer[:a].copy_(stage[:a])
        if b < total:
            buffer[b:total].copy_(stage[b:total])


def exchange_kind(numel=None):
    """What carries the offline exchange in this process (for bench lines and logs)."""
    if not is_distributed():
        return 'none (one rank)'
    st = PeerStaging._current
    if st is not None:
        return f'peer memory ({st.kind}) + signal-pad barrier'
    why = f' [{PeerStaging._disabled}]' if PeerStaging._disabled else ''
    return 'NCCL all_gather_into_tensor' + why


def exchange_regions(buffer, starts, handle=None):
    """All-gather of disjoint contiguous regions of ``buffer`` (rank ``r`` owns ``buffer[starts[r]:starts[r+1]]``).

    With a library ``handle`` on an NCCL process group the regions travel over peer memory (``PeerStaging``: one
    ``peer_push_kernel`` + a signal-pad barrier, no NCCL launch).  Otherwise -- equal-length regions (what
    ``region_layout`` produces): one in-place ``all_gather_into_tensor``, a single NCCL collective over NVLink /
    NVSwitch, launch-latency-bound at these sizes (tens of MB), where eight serialised broadcasts cost eight launches.
    Unequal regions (foreign layouts) fall back to one broadcast per rank."""
    if not is_distributed():
        return
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    aligned = buffer.data_ptr() % 16 == 0 and all(int(x) % 2 == 0 for x in starts)     # 16-byte stores in peer_push_kernel
    if handle is not None and buffer.is_cuda and aligned and int(starts[world]) <= buffer.numel():
        staging = PeerStaging.acquire(int(starts[world]))
        if staging is not None:
            staging.exchange(handle, buffer, starts)
            return
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
