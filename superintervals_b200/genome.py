"""Contig-partitioned container (SURVEY.md 8e mode B): one superset index per contig.

The reference has no multi-contig type of its own: its callers keep one IntervalMap per
chromosome and route every query to the map of its contig
(reference examples/bed-intersect-si.rs:100-123; test/bench.cpp:67-102 handles one
chromosome). ``GenomeIndex`` is that container for one process per GPU: contigs are
assigned to ranks by longest-processing-time on (N_c + Q_c), a rank builds and queries
only the contigs it owns, and the only exchange between ranks is the per-contig hit
totals (one small all_gather over NCCL / gloo) from which every rank derives the base
offset of each contig's segment in the global, contig-major CSR.

PyTorch is plumbing here (device memory, the routing sort, torch.distributed); every
query kernel is the library's.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .sharding import assign_contigs

__all__ = ["GenomeIndex", "route_by_contig"]


def route_by_contig(contig_ids, n_contigs):
    """Stable grouping of a mixed query batch by contig id (host side, numpy).
    Returns (order, bounds): queries order[bounds[c]:bounds[c+1]] belong to contig c, in
    their original relative order."""
    cid = np.ascontiguousarray(contig_ids)
    if cid.size and (cid.min() < 0 or cid.max() >= n_contigs):
        raise ValueError("contig id out of range")
    order = np.argsort(cid, kind="stable")
    bounds = np.zeros(n_contigs + 1, np.int64)
    np.cumsum(np.bincount(cid, minlength=n_contigs), out=bounds[1:])
    return order, bounds


class GenomeIndex:
    """Per-contig indexes of the contigs this rank owns."""

    def __init__(self, names, n_intervals, n_queries=None, rank=None, world=None):
        self.names = list(names)
        nc = len(self.names)
        if len(n_intervals) != nc:
            raise ValueError("one interval count per contig")
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank, self.world = int(rank), int(world)
        nq = n_queries if n_queries is not None else [0] * nc
        self.owner = assign_contigs(n_intervals, nq, self.world)      # owner rank of every contig
        self._ix = {}                                                  # contig -> DeviceIndex (owned contigs only)
        self.hits = np.zeros(nc, np.int64)                             # hit totals of the last count per owned contig

    def owns(self, c):
        return int(self.owner[c]) == self.rank

    @property
    def owned(self):
        return [c for c in range(len(self.names)) if self.owns(c)]

    def build_contig(self, c, starts, ends, values=None):
        """build() of contig c from device int32 tensors; only the owning rank may call it."""
        from .device import DeviceIndex
        if not self.owns(c):
            raise ValueError(f"rank {self.rank} does not own contig {self.names[c]} (owner {int(self.owner[c])})")
        self._ix[c] = DeviceIndex().build(starts, ends, values)
        return self._ix[c]

    def index(self, c):
        return self._ix[c]

    def count_contig(self, c, qs, qe, out=None, order=0):
        """Counts for queries that all lie on contig c (device tensors)."""
        return self._ix[c].count(qs, qe, out=out, order=order)

    def count(self, contig_ids, qs, qe):
        """Counts for a mixed host batch (numpy: contig id, start, end per query). Queries of
        contigs this rank does not own get 0 here -- their owner answers them; summing the
        vectors of all ranks (or indexing by owner) gives the full answer."""
        order, bounds = route_by_contig(contig_ids, len(self.names))
        qs = np.ascontiguousarray(qs, np.int32)
        qe = np.ascontiguousarray(qe, np.int32)
        out = np.zeros(qs.shape[0], np.uint32)
        self.hits[:] = 0
        for c in self.owned:
            lo, hi = int(bounds[c]), int(bounds[c + 1])
            if hi == lo or c not in self._ix:
                continue
            sel = order[lo:hi]
            d = self._ix[c].count(torch.from_numpy(qs[sel]).cuda(), torch.from_numpy(qe[sel]).cuda())
            cnt = d.cpu().numpy().astype(np.uint32)
            out[sel] = cnt
            self.hits[c] = int(cnt.astype(np.int64).sum())
        return out

    def csr_bases(self, hits=None, device=None, group=None):
        """All-gather the per-contig hit totals. Returns (bases, totals) int64[n_contigs]:
        contig c's values occupy [bases[c], bases[c] + totals[c]) of the global contig-major CSR.
        Each contig's total is contributed by its owner (other ranks hold 0 for it)."""
        mine = np.asarray(self.hits if hits is None else hits, np.int64).copy()
        mine[np.asarray(self.owner) != self.rank] = 0
        if dist.is_available() and dist.is_initialized() and self.world > 1:
            t = torch.from_numpy(mine)
            if device is not None:
                t = t.to(device)
            parts = [torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t, group=group)
            totals = torch.stack(parts).sum(0).cpu().numpy()
        else:
            totals = mine
        bases = np.zeros_like(totals)
        np.cumsum(totals[:-1], out=bases[1:])
        return bases, totals
