"""Multi-GPU sharding of the batch query path: one process per GPU.

Queries are independent and the index is read-only after build(), so the path shards
with NO data-path collective:

  mode A (replicated index)   every rank builds/holds the whole index and answers a
                              contiguous slice of the query batch;
  mode B (contig-partitioned) one index per contig (the reference's own convention,
                              examples/bed-intersect-si.rs:100-123), contigs assigned to
                              ranks by longest-processing-time on (N_c + Q_c).

The only exchange is the CSR bookkeeping: each rank's total hit count is all-gathered
so every rank knows the base offset of its value segment in the global CSR (a rank's
segment is contiguous because its query range is). Backend-agnostic: NCCL over
NVLink on the B200 box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .workloads import lpt_assign, shard_range

__all__ = ["shard_range", "lpt_assign", "csr_shard_bases", "gather_counts", "assign_contigs", "PeerGathered", "PeerBatch"]


def csr_shard_bases(local_total_hits: int, device=None, group=None):
    """All-gather per-rank hit totals. Returns (bases[world], totals[world]) as Python ints:
    rank r's values occupy [bases[r], bases[r] + totals[r]) of the global CSR."""
    if not (dist.is_available() and dist.is_initialized()):
        return [0], [int(local_total_hits)]
    world = dist.get_world_size(group)
    mine = torch.tensor([int(local_total_hits)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    totals = [int(t.item()) for t in out]
    bases = [0] * world
    for r in range(1, world):
        bases[r] = bases[r - 1] + totals[r - 1]
    return bases, totals


def gather_counts(local_counts: torch.Tensor, n_total: int, group=None):
    """Mode A: assemble the full per-query count vector on every rank from contiguous
    slices (shard_range order). Slices may differ by one element, so pad to the max."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_counts
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(width, dtype=local_counts.dtype, device=local_counts.device)
    buf[: local_counts.numel()] = local_counts
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])


def assign_contigs(n_intervals, n_queries, world):
    """Mode B: owner rank of every contig, LPT on (N_c + Q_c)."""
    cost = np.asarray(n_intervals, np.float64) + np.asarray(n_queries, np.float64)
    return lpt_assign(cost, world)


class PeerGathered:
    """Count vectors gathered WITHOUT a collective: two int32 arrays of `world * per` entries (double buffer) plus one
    flag word per rank on every GPU, each GPU's block mapped into every other rank's address space (CUDA IPC over
    NVLink peer access; siIpcAlloc / siIpcOpen). Rank r's count kernel stores its slot [r * per, (r + 1) * per) into
    EVERY copy itself (siCountFanoutDevice), then a one-warp kernel signals and awaits the peers' flags
    (siPeerBarrierDevice): "count, then all_gather of the counts" is one count kernel + one barrier kernel.
    Steps alternate between the two arrays, so a fast rank's stores of step k+1 never land in an array a slower rank
    is still reading for step k (it cannot pass step k+1's barrier, hence cannot start step k+2, before that rank has
    finished step k). Collective: every rank of the group constructs it; close() before the process group goes away."""

    def __init__(self, per, world, rank, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        if world - 1 > 15:
            raise ValueError("at most 16 ranks")
        self._L, self._lib, self._C = _lib.lib(), _lib, C
        self.per, self.world, self.rank = int(per), int(world), int(rank)
        self.n = self.per * self.world
        self._flag_off = 2 * self.n * 4                     # bytes: [array 0][array 1][flags: world words][timed_out word]
        total = self._flag_off + 4 * (world + 1)
        total = (total + 255) & ~255
        def agreed(ok, what):       # every rank learns whether every rank succeeded: a failure must not leave the others in a collective
            t = torch.tensor([1.0 if ok else 0.0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            if t.item() < 1.0:
                _lib.lib().si_b200_clear_error()
                raise RuntimeError(f"PeerGathered: {what} failed on at least one rank")

        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        rc = self._L.siIpcAlloc(total, C.byref(ptr), handle)
        self._own = ptr.value if rc == 0 else None
        self._peers = {}
        try:
            agreed(rc == 0, "siIpcAlloc (cudaMalloc + cudaIpcGetMemHandle)")
        except RuntimeError:
            if self._own:
                self._L.siIpcFree(self._own)
            raise
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        allh = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = allh.cpu().numpy()
        opened = True
        for r in range(world):
            if r == rank:
                continue
            p = C.c_void_p()
            hb = (C.c_ubyte * 64)(*allh[r * 64:(r + 1) * 64].tolist())
            if self._L.siIpcOpen(hb, C.byref(p)) != 0:
                opened = False
                break
            self._peers[r] = p.value
        try:
            agreed(opened, "siIpcOpen (cudaIpcOpenMemHandle: peer access between the GPUs)")
        except RuntimeError:
            for p in self._peers.values():
                self._L.siIpcClose(p)
            self._L.siIpcFree(self._own)
            raise

        class _Raw:   # __cuda_array_interface__ view of the cudaMalloc'ed block
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (total // 4,), "typestr": "<i4", "data": (self._own, False), "version": 3, "strides": None}
        self._raw = raw
        self._all = torch.as_tensor(raw, device="cuda")
        self._all.zero_()
        self.arrays = [self._all[:self.n], self._all[self.n:2 * self.n]]
        self._timed_out = self._all[2 * self.n + world: 2 * self.n + world + 1]
        order = sorted(self._peers)
        self._signal = (C.c_void_p * max(1, len(order)))(*[C.c_void_p(self._peers[r] + self._flag_off + 4 * rank) for r in order])
        self._wait = (C.c_void_p * max(1, len(order)))(*[C.c_void_p(self._own + self._flag_off + 4 * r) for r in order])
        self._order = order
        self.step = 0
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def current(self):
        """(this step's gathered array, this rank's slot in it, addresses of this slot in the peers' copies)"""
        b = self.step & 1
        arr = self.arrays[b]
        slot = arr[self.rank * self.per:(self.rank + 1) * self.per]
        ptrs = [self._peers[r] + 4 * (b * self.n + self.rank * self.per) for r in self._order]
        return arr, slot, ptrs

    def barrier(self):
        """Stream-ordered: returns (on the stream) once every peer's stores of this step have landed; advances the step."""
        import torch
        self.step += 1
        self._L.siPeerBarrierDevice(self._signal, self._wait, len(self._order), self.step, self._timed_out.data_ptr(),
                                    self._C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check("siPeerBarrierDevice")

    def timed_out(self):
        return bool(int(self._timed_out.item()))

    def close(self):
        import torch
        torch.cuda.synchronize()
        for p in self._peers.values():
            self._L.siIpcClose(p)
        self._peers = {}
        self._all = self.arrays = self._timed_out = None
        if self._own:
            self._L.siIpcFree(self._own)
            self._own = None


class PeerBatch:
    """A mixed query batch that stays where it was produced: every rank owns one slice (contig id, start, end per query, and
    the counts that come back) in a device block every other rank has mapped (CUDA IPC over NVLink peer access; siIpcAlloc /
    siIpcOpen). GenomeIndex.count_mixed_peer() lets each rank walk ALL slices in place and answer the queries of the contigs
    it indexes, storing the counts straight into the slice they belong to -- mode B of SURVEY 8e without the all-to-all
    dispatch and combine. Layout of a block (int32 words): [contig cap][qs cap][qe cap][counts cap][flags: world][timed_out].
    Collective: every rank of the group constructs it with the same `cap` (queries per slice at most); close() before the
    process group goes away."""

    def __init__(self, cap, world, rank, group=None):
        import ctypes as C
        from . import _lib
        if world - 1 > 15:
            raise ValueError("at most 16 ranks")
        self._L, self._lib, self._C = _lib.lib(), _lib, C
        self.cap = (int(cap) + 63) & ~63
        self.world, self.rank, self.group = int(world), int(rank), group
        caps = torch.tensor([float(self.cap), -float(self.cap)], dtype=torch.float64, device="cuda")
        dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)
        if int(caps[0].item()) != -int(caps[1].item()):   # the peers' blocks are addressed with this rank's cap
            raise ValueError("PeerBatch: every rank must pass the same capacity")
        self._flag_off = 4 * self.cap * 4
        total = (self._flag_off + 4 * (world + 1) + 255) & ~255

        def agreed(ok, what):
            t = torch.tensor([1.0 if ok else 0.0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            if t.item() < 1.0:
                _lib.lib().si_b200_clear_error()
                raise RuntimeError(f"PeerBatch: {what} failed on at least one rank")

        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        rc = self._L.siIpcAlloc(total, C.byref(ptr), handle)
        self._own = ptr.value if rc == 0 else None
        self._peers = {}
        try:
            agreed(rc == 0, "siIpcAlloc (cudaMalloc + cudaIpcGetMemHandle)")
        except RuntimeError:
            if self._own:
                self._L.siIpcFree(self._own)
            raise
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        allh = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = allh.cpu().numpy()
        opened = True
        for r in range(world):
            if r == rank:
                continue
            p = C.c_void_p()
            hb = (C.c_ubyte * 64)(*allh[r * 64:(r + 1) * 64].tolist())
            if self._L.siIpcOpen(hb, C.byref(p)) != 0:
                opened = False
                break
            self._peers[r] = p.value
        try:
            agreed(opened, "siIpcOpen (cudaIpcOpenMemHandle: peer access between the GPUs)")
        except RuntimeError:
            for p in self._peers.values():
                self._L.siIpcClose(p)
            self._L.siIpcFree(self._own)
            raise

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (total // 4,), "typestr": "<i4", "data": (self._own, False), "version": 3, "strides": None}
        self._raw = raw
        self._all = torch.as_tensor(raw, device="cuda")
        self._all.zero_()
        c = self.cap
        self.contig, self.qs, self.qe, self.counts = (self._all[k * c:(k + 1) * c] for k in range(4))
        self._timed_out = self._all[4 * c + world: 4 * c + world + 1]
        order = sorted(self._peers)
        self._signal = (C.c_void_p * max(1, len(order)))(*[C.c_void_p(self._peers[r] + self._flag_off + 4 * rank) for r in order])
        self._wait = (C.c_void_p * max(1, len(order)))(*[C.c_void_p(self._own + self._flag_off + 4 * r) for r in order])
        self._order = order
        self.step = 0
        self.lengths = [0] * world
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def base(self, r):
        """device address of rank r's block in THIS process's address space"""
        return self._own if r == self.rank else self._peers[r]

    def set_length(self, n):
        """this rank's slice holds n queries (contig[:n], qs[:n], qe[:n]); collective: every rank learns every length"""
        if n > self.cap:
            raise ValueError("slice longer than the batch's capacity")
        t = torch.tensor([int(n)], dtype=torch.int64, device="cuda")
        out = torch.empty(self.world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(out, t, group=self.group)
        self.lengths = [int(x) for x in out.cpu().tolist()]
        return self.lengths

    def barrier(self):
        """Stream-ordered barrier between the ranks (siPeerBarrierDevice): flags through peer memory, no collective call."""
        self.step += 1
        self._L.siPeerBarrierDevice(self._signal, self._wait, len(self._order), self.step, self._timed_out.data_ptr(),
                                    self._C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check("siPeerBarrierDevice")

    def timed_out(self):
        return bool(int(self._timed_out.item()))

    def close(self):
        torch.cuda.synchronize()
        for p in self._peers.values():
            self._L.siIpcClose(p)
        self._peers = {}
        self._all = self.contig = self.qs = self.qe = self.counts = self._timed_out = None
        if self._own:
            self._L.siIpcFree(self._own)
            self._own = None
