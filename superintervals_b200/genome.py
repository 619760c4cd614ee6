"""Contig-partitioned container (SURVEY.md 8e mode B): one superset index per contig.

The reference has no multi-contig type of its own: its callers keep one IntervalMap per
chromosome and route every query to the map of its contig
(reference examples/bed-intersect-si.rs:100-123; test/bench.cpp:67-102 handles one
chromosome). ``GenomeIndex`` is that container for one process per GPU: contigs are
assigned to ranks by longest-processing-time on (N_c + Q_c), a rank builds and queries
only the contigs it owns.

A mixed batch (contig id, start, end per query) is answered in three device steps:

  route     group the batch by contig ON THE DEVICE (siRouteByContigDevice: the build's
            radix sort over the contig ids, the query columns gathered through it);
  dispatch  with several ranks each rank holds a slice of the batch: every query travels to
            the rank that owns its contig -- one all-to-all over NVLink (12 B per query:
            contig, start, end), the counts travel back by the inverse all-to-all (4 B per
            query). With one rank nothing moves;
  count     contig c's queries are one contiguous device range for contig c's index.

When this rank holds every contig (one rank, or the indexes replicated) nothing is routed at all: count_mixed() is ONE
launch of the mixed-batch kernel over the batch in the caller's order (siCountMixedDevice: per-contig rank-cell descriptors
in shared memory), and with the contigs partitioned the owner counts what arrived in one such launch.
search_values_mixed() returns the contig-major CSR of a mixed batch (count -> scan -> per-contig fill).
The per-contig hit totals are all-gathered into the base offsets of a global contig-major CSR.
PyTorch is plumbing here (device memory, torch.distributed); routing, counting and the scatter
back are the library's kernels. The route / count steps are injectable so that the exchange logic
runs under gloo on CPU in the tests (tests/test_genome.py) -- the defaults are the CUDA library and
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from .sharding import assign_contigs

__all__ = ["GenomeIndex", "route_by_contig"]


def route_by_contig(contig_ids, n_contigs):
    """Stable grouping of a mixed query batch by contig id (host side, numpy): the definition the device
    routing is tested against. Returns (order, bounds): queries order[bounds[c]:bounds[c+1]] belong to
    contig c, in their original relative order."""
    cid = np.ascontiguousarray(contig_ids)
    if cid.size and (cid.min() < 0 or cid.max() >= n_contigs):
        raise ValueError("contig id out of range")
    order = np.argsort(cid, kind="stable")
    bounds = np.zeros(n_contigs + 1, np.int64)
    np.cumsum(np.bincount(cid, minlength=n_contigs), out=bounds[1:])
    return order, bounds


class _CudaRouter:
    """route / scatter through libsuperintervals_b200.so (siRouteByContigDevice, siScatterCountsDevice)."""

    def __init__(self):
        from . import _lib
        from .device import DeviceIndex
        self._lib = _lib
        self._L = _lib.lib()
        self._scratch = DeviceIndex()          # lends its sort scratch; never built

    def route(self, key, qs, qe, n_keys):
        n = key.numel()
        gs, ge = torch.empty_like(qs), torch.empty_like(qe)
        perm = torch.empty(n, dtype=torch.int32, device=key.device)
        off = (C.c_size_t * (n_keys + 1))()
        self._L.siRouteByContigDevice(self._scratch._ix, key.data_ptr(), qs.data_ptr(), qe.data_ptr(), n, n_keys,
                                      gs.data_ptr(), ge.data_ptr(), perm.data_ptr(), off,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check("siRouteByContigDevice")
        return gs, ge, perm, np.frombuffer(off, dtype=np.uint64).astype(np.int64)

    def count_mixed(self, indexes, contig, qs, qe, out):
        """out[i] = count of query i on indexes[contig[i]] (None: 0), caller's order, ONE launch, nothing routed
        (siCountMixedDevice). Returns the per-entry hit totals (int64 numpy), or None when an index cannot answer
        from rank cells: the caller routes by contig instead."""
        n_c = len(indexes)
        arr = (C.c_void_p * n_c)(*[ix._ix if ix is not None else None for ix in indexes])
        totals = torch.empty(n_c, dtype=torch.int64, device=contig.device)
        rc = self._L.siCountMixedDevice(arr, n_c, contig.data_ptr(), qs.data_ptr(), qe.data_ptr(), contig.numel(), out.data_ptr(),
                                        totals.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc == -2:                                        # SI_MIXED_UNSUPPORTED
            return None
        self._lib.check("siCountMixedDevice")
        return totals.cpu().numpy()

    def scatter(self, counts, perm, out):
        self._L.siScatterCountsDevice(self._scratch._ix, counts.data_ptr(), perm.data_ptr(), counts.numel(), out.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check("siScatterCountsDevice")
        return out


def _lib_opt_pair_cells():
    from ._lib import OPT_PAIR_CELLS
    return OPT_PAIR_CELLS


class GenomeIndex:
    """Per-contig indexes of the contigs this rank owns."""

    def __init__(self, names, n_intervals, n_queries=None, rank=None, world=None, router=None, group=None, pair_cells=None):
        self.names = list(names)
        nc = len(self.names)
        self._n_intervals = [int(x) for x in n_intervals]
        self._pair_cells = pair_cells      # SI_OPT_PAIR_CELLS of every contig's index; None: decided from the genome's size
        if len(n_intervals) != nc:
            raise ValueError("one interval count per contig")
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank, self.world = int(rank), int(world)
        self.group = group
        nq = n_queries if n_queries is not None else [0] * nc
        self.owner = np.asarray(assign_contigs(n_intervals, nq, self.world), np.int64)   # owner rank of every contig
        # slots: the contigs in (owner, contig) order -- routing by slot groups a batch by destination rank first
        self.slot_contig = np.lexsort((np.arange(nc), self.owner)).astype(np.int64)
        self.slot_of = np.empty(nc, np.int64)
        self.slot_of[self.slot_contig] = np.arange(nc)
        self.slot_bounds = np.searchsorted(self.owner[self.slot_contig], np.arange(self.world + 1))  # rank r owns slots [b[r], b[r+1])
        self._ix = {}                                                  # contig -> DeviceIndex (owned contigs only)
        self.hits = np.zeros(nc, np.int64)                             # hit totals of the last count per owned contig
        self._router = router
        self._lut = None
        self.last_exchange = {"dispatch_bytes": 0, "combine_bytes": 0}

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
        ix = DeviceIndex()
        ix.set_option(_lib_opt_pair_cells(), self._pair_mode())
        self._ix[c] = ix.build(starts, ends, values)
        return self._ix[c]

    def _pair_mode(self):
        """One index per contig: a mixed batch gathers from the rank cells of ALL owned contigs at once, so whether
        they fit L2 is a property of the genome, not of one contig (~6 B of rank cells per interval against 3/4 of
        the 126 MB L2). Beyond that every contig also gets pair cells: one sector per short query from HBM."""
        if self._pair_cells is not None:
            return int(self._pair_cells)
        owned = sum(n for c, n in enumerate(self._n_intervals) if self.owns(c))
        return 2 if owned * 6 > 96_000_000 else 1

    def index(self, c):
        return self._ix[c]

    def count_contig(self, c, qs, qe, out=None, order=0):
        """Counts for queries that all lie on contig c (device tensors)."""
        return self._ix[c].count(qs, qe, out=out, order=order)

    # ---- mixed batches --------------------------------------------------------------------------------
    def _route(self):
        if self._router is None:
            self._router = _CudaRouter()
        return self._router

    def _count_slots(self, gs, ge, off, slots, count_fn):
        """counts of routed queries: slot k's queries are gs/ge[off[k]:off[k+1]] (k indexes `slots`)."""
        counts = torch.zeros(gs.numel(), dtype=torch.int32, device=gs.device)
        for k, slot in enumerate(slots):
            lo, hi = int(off[k]), int(off[k + 1])
            c = int(self.slot_contig[slot])
            if hi == lo:
                continue
            if count_fn is not None:
                count_fn(c, gs[lo:hi], ge[lo:hi], counts[lo:hi])
            elif c in self._ix:
                self._ix[c].count(gs[lo:hi], ge[lo:hi], out=counts[lo:hi])
        return counts

    def count_mixed(self, contig_ids, qs, qe, count_fn=None):
        """Counts, in the caller's order, for THIS RANK'S slice of a mixed batch (int32 tensors on this rank's device:
        contig id, start, end per query). Every query is answered by the rank that owns its contig: route by
        destination, all-to-all dispatch, count, all-to-all combine. Collective: every rank of the group must call it."""
        n = contig_ids.numel()
        dev = contig_ids.device
        nc = len(self.names)
        R = self._route()
        if self._lut is None or self._lut.device != dev:
            self._lut = torch.from_numpy(self.slot_of.astype(np.int32)).to(dev)
        if n and (int(contig_ids.min()) < 0 or int(contig_ids.max()) >= nc):
            raise ValueError("contig id out of range")
        direct = count_fn is None and hasattr(R, "count_mixed")    # the one-launch kernel (no grouping by contig)
        table = [self._ix.get(c) for c in range(nc)] if direct else None
        if self.world == 1:                                            # one rank owns every contig: nothing moves
            self.last_exchange = {"dispatch_bytes": 0, "combine_bytes": 0}
            if direct:
                out = torch.empty(n, dtype=torch.int32, device=dev)
                tot = R.count_mixed(table, contig_ids, qs, qe, out)
                if tot is not None:
                    self.hits[:] = tot
                    return out
            gs, ge, perm, off = R.route(contig_ids, qs, qe, nc)
            routed = self._count_slots(gs, ge, off, list(range(nc)), count_fn)
            self._sum_hits(routed, off, list(range(nc)))
            out = torch.empty(n, dtype=torch.int32, device=dev)
            return R.scatter(routed, perm, out) if n else out
        slot = self._lut[contig_ids.long()]
        # ---- several ranks: dispatch by owner
        gs, ge, perm, off = R.route(slot, qs, qe, nc)                 # slot-major = destination-major
        gslot = slot[perm.long()] if n else slot                       # routed slot ids travel with the queries
        send = np.array([off[self.slot_bounds[r + 1]] - off[self.slot_bounds[r]] for r in range(self.world)], np.int64)
        t_send = torch.from_numpy(send).to(dev)
        t_recv = torch.empty_like(t_send)
        dist.all_to_all_single(t_recv, t_send, group=self.group)
        recv = t_recv.cpu().numpy()
        m = int(recv.sum())
        rbuf = [torch.empty(m, dtype=torch.int32, device=dev) for _ in range(3)]
        for dst, src in zip(rbuf, (gslot.to(torch.int32), gs, ge)):
            dist.all_to_all_single(dst, src.contiguous(), output_split_sizes=recv.tolist(), input_split_sizes=send.tolist(),
                                   group=self.group)
        self.last_exchange = {"dispatch_bytes": 12 * int(send.sum() - send[self.rank]), "combine_bytes": 4 * int(send.sum() - send[self.rank])}
        my0, my1 = int(self.slot_bounds[self.rank]), int(self.slot_bounds[self.rank + 1])
        back = torch.empty(m, dtype=torch.int32, device=dev)
        answered = False
        if direct:
            # what arrived is in source-major order and stays there: one launch over this rank's slots
            slot_table = [None] * nc
            for sl in range(my0, my1):
                slot_table[sl] = self._ix.get(int(self.slot_contig[sl]))
            tot = R.count_mixed(slot_table, rbuf[0], rbuf[1], rbuf[2], back)
            answered = tot is not None
            if answered:
                self.hits[:] = 0
                self.hits[self.slot_contig] = tot
        if not answered:
            # group what arrived by my slots, count per contig, back to arrival order
            local = rbuf[0] - my0
            hs, he, hperm, hoff = R.route(local, rbuf[1], rbuf[2], max(1, my1 - my0))
            hcounts = self._count_slots(hs, he, hoff, list(range(my0, my1)), count_fn)
            self._sum_hits(hcounts, hoff, list(range(my0, my1)))
            if m:
                R.scatter(hcounts, hperm, back)                        # arrival order = source-major
        routed = torch.empty(n, dtype=torch.int32, device=dev)
        dist.all_to_all_single(routed, back, output_split_sizes=send.tolist(), input_split_sizes=recv.tolist(), group=self.group)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        return R.scatter(routed, perm, out) if n else out

    def count_mixed_peer(self, batch):
        """Mode B without a dispatch: `batch` is a sharding.PeerBatch whose slices the ranks have filled (set_length done).
        Every rank walks all slices in place -- the remote ones over NVLink -- answers the queries of the contigs it holds
        an index for and stores each count into the slice it belongs to (siCountMixedPeerDevice); two flag barriers through
        peer memory bracket the walk. Returns this rank's counts (a view of batch.counts, the caller's order), or None when
        the kernel cannot run here (an index without rank cells, or a rank that owns no indexed contig): route instead.
        Collective. self.hits receives the totals of the owned contigs over the WHOLE batch; last_exchange the NVLink bytes."""
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        nc, W = len(self.names), self.world
        arr = (C.c_void_p * nc)(*[(self._ix[c]._ix if c in self._ix else None) for c in range(nc)])
        foreign = (C.c_ubyte * nc)(*[1 if (not self.owns(c) and self._n_intervals[c] > 0) else 0 for c in range(nc)])
        cap = batch.cap

        def ptrs(k):
            return (C.c_void_p * W)(*[C.c_void_p(batch.base(r) + 4 * k * cap) for r in range(W)])
        lens = (C.c_size_t * W)(*batch.lengths)
        totals = torch.zeros(nc, dtype=torch.int64, device=batch.counts.device)
        batch.barrier()                       # every slice is complete before a peer reads it
        rc = L.siCountMixedPeerDevice(arr, nc, foreign, W, self.rank, ptrs(0), ptrs(1), ptrs(2), lens, ptrs(3), totals.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
        batch.barrier()                       # every rank's stores have landed before the counts are read
        if rc == -2:                          # SI_MIXED_UNSUPPORTED
            return None
        _lib.check("siCountMixedPeerDevice")
        self.hits[:] = totals.cpu().numpy()
        away = sum(n for r, n in enumerate(batch.lengths) if r != self.rank)
        self.last_exchange = {"contig_id_bytes_read_remote": 4 * away, "note": "plus 8 B read and 4 B written per remote query of an owned contig"}
        return batch.counts[:batch.lengths[self.rank]]

    def _sum_hits(self, counts, off, slots):
        self.hits[:] = 0
        for k, slot in enumerate(slots):
            lo, hi = int(off[k]), int(off[k + 1])
            if hi > lo:
                self.hits[int(self.slot_contig[slot])] = int(counts[lo:hi].to(torch.int64).sum().item())

    def count(self, contig_ids, qs, qe):
        """Counts for a mixed HOST batch (numpy: contig id, start, end per query), routed on the device. Queries of
        contigs this rank does not own get 0 here -- their owner answers them; summing the vectors of all
        ranks gives the full answer. (No exchange: every rank sees the whole batch. count_mixed() is the
        sharded form.)"""
        cid = torch.from_numpy(np.ascontiguousarray(contig_ids, np.int32)).cuda()
        dqs = torch.from_numpy(np.ascontiguousarray(qs, np.int32)).cuda()
        dqe = torch.from_numpy(np.ascontiguousarray(qe, np.int32)).cuda()
        n, nc = cid.numel(), len(self.names)
        if n and (int(cid.min()) < 0 or int(cid.max()) >= nc):
            raise ValueError("contig id out of range")
        R = self._route()
        gs, ge, perm, off = R.route(cid, dqs, dqe, nc)
        routed = torch.zeros(n, dtype=torch.int32, device=cid.device)
        self.hits[:] = 0
        for c in self.owned:
            lo, hi = int(off[c]), int(off[c + 1])
            if hi == lo or c not in self._ix:
                continue
            self._ix[c].count(gs[lo:hi], ge[lo:hi], out=routed[lo:hi])
            self.hits[c] = int(routed[lo:hi].to(torch.int64).sum().item())
        out = torch.empty(n, dtype=torch.int32, device=cid.device)
        if n:
            R.scatter(routed, perm, out)
        return out.cpu().numpy().astype(np.uint32)

    def search_values_mixed(self, contig_ids, qs, qe, what=None):
        """search_values for a mixed batch on the contigs THIS rank holds (every contig when the indexes are replicated or
        world == 1): the batch is grouped by contig on the device and each contig's queries are one count -> scan -> fill
        on that contig's index. Returns (perm, offsets, values): routed query k is the caller's query perm[k] (contig-major,
        stable), its hits are values[offsets[k]:offsets[k + 1]] in the reference's descending position order, and contig
        c's segment starts at the base csr_bases() reports -- the global contig-major CSR of SURVEY 8e. Queries of contigs
        without an index here get empty lists."""
        from .device import FILL_VALUES
        what = FILL_VALUES if what is None else what
        n, nc, dev = contig_ids.numel(), len(self.names), contig_ids.device
        if n and (int(contig_ids.min()) < 0 or int(contig_ids.max()) >= nc):
            raise ValueError("contig id out of range")
        R = self._route()
        gs, ge, perm, off = R.route(contig_ids, qs, qe, nc)
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        for c in range(nc):
            lo, hi = int(off[c]), int(off[c + 1])
            if hi > lo and c in self._ix:
                self._ix[c].count(gs[lo:hi], ge[lo:hi], out=counts[lo:hi])
        any_ix = next(iter(self._ix.values()), None)
        if any_ix is None or n == 0:
            return perm, torch.zeros(n + 1, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int32, device=dev)
        offsets = any_ix.scan(counts)
        bounds = offsets[torch.as_tensor(off, device=dev)].cpu().numpy()      # where each contig's values begin
        values = torch.empty(int(bounds[-1]), dtype=torch.int32, device=dev)
        self.hits[:] = 0
        for c in range(nc):
            lo, hi = int(off[c]), int(off[c + 1])
            if hi == lo or c not in self._ix:
                continue
            self.hits[c] = int(bounds[c + 1] - bounds[c])
            # the contig's own count -> scan -> fill (fresh, aligned buffers), its values copied into its segment
            _, v_c = self._ix[c].search(gs[lo:hi].contiguous(), ge[lo:hi].contiguous(), what=what)
            values[int(bounds[c]):int(bounds[c + 1])] = v_c
        return perm, offsets, values

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
