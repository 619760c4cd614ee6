"""Contig-partitioned mode (SURVEY 8e mode B): routing and CSR bookkeeping on CPU (gloo, world
size 2), per-contig parity against the oracle on the GPU."""
import multiprocessing as mp
import socket

import numpy as np
import pytest

from superintervals_b200 import workloads as W


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_route_by_contig_is_a_stable_grouping():
    from superintervals_b200.genome import route_by_contig
    rng = np.random.default_rng(0)
    cid = rng.integers(0, 5, 1000)
    order, bounds = route_by_contig(cid, 5)
    assert bounds[0] == 0 and bounds[-1] == 1000
    for c in range(5):
        sel = order[bounds[c]:bounds[c + 1]]
        assert (cid[sel] == c).all() and (np.diff(sel) > 0).all()      # stable: original relative order
    with pytest.raises(ValueError):
        route_by_contig(np.array([0, 7]), 5)


def test_contigs_are_owned_exactly_once_and_balanced():
    from superintervals_b200.genome import GenomeIndex
    parts = W.config4_partition(1_000_000, 10_000_000)
    n_c = [p[0] for p in parts]; q_c = [p[1] for p in parts]
    for world in (1, 2, 4, 8):
        gs = [GenomeIndex([f"chr{i}" for i in range(24)], n_c, q_c, rank=r, world=world) for r in range(world)]
        owned = sorted(c for g in gs for c in g.owned)
        assert owned == list(range(24))
        load = [sum(n_c[c] + q_c[c] for c in g.owned) for g in gs]
        assert max(load) <= 1.25 * (sum(load) / world)                   # LPT keeps ranks within 25 % of the mean


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from superintervals_b200.genome import GenomeIndex
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_c = [5, 50, 20, 7, 30, 1]
    g = GenomeIndex(list("abcdef"), n_c, n_c)
    hits = np.arange(1, 7, dtype=np.int64) * 10        # what each contig WOULD total; only owners contribute
    bases, totals = g.csr_bases(hits)
    ok = np.array_equal(totals, hits) and np.array_equal(bases, np.concatenate([[0], np.cumsum(hits)[:-1]]))
    ok &= sorted(g.owned) == [c for c in range(6) if g.owner[c] == rank]
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world_size_2_contig_csr_bases_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs: p.start()
        for p in procs: p.join(120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}


class _HostRouter:
    """The routing step restated with numpy (route_by_contig) for the gloo test of the exchange logic."""

    def route(self, key, qs, qe, n_keys):
        import torch
        from superintervals_b200.genome import route_by_contig
        order, bounds = route_by_contig(key.numpy().astype(np.int64), n_keys)
        o = torch.from_numpy(order)
        return qs[o].contiguous(), qe[o].contiguous(), o.to(torch.int32), bounds

    def scatter(self, counts, perm, out):
        out[perm.long()] = counts
        return out


def _mixed_case(n_contigs=6, seed=3):
    rng = np.random.default_rng(seed)
    data = []
    for c in range(n_contigs):
        n = int(rng.integers(50, 400))
        s = rng.integers(0, 20_000, n).astype(np.int32)
        e = (s + rng.integers(0, 900, n)).astype(np.int32)
        data.append((s, e))
    nq = 5000
    cid = rng.integers(0, n_contigs, nq).astype(np.int32)
    qs = rng.integers(0, 20_000, nq).astype(np.int32)
    qe = (qs + rng.integers(0, 1500, nq)).astype(np.int32)
    return data, cid, qs, qe


def _mixed_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from superintervals_b200.genome import GenomeIndex
    from superintervals_b200.workloads import shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    data, cid, qs, qe = _mixed_case()
    g = GenomeIndex([f"c{i}" for i in range(len(data))], [d[0].size for d in data], None, router=_HostRouter())
    oracles = {c: Oracle(*data[c]) for c in g.owned}
    asked = []

    def count_fn(c, a, b, dst):                      # only the owner may be asked about a contig
        asked.append(c)
        dst.copy_(torch.from_numpy(oracles[c].count_batch(a.numpy(), b.numpy()).astype(np.int32)))

    lo, hi = shard_range(cid.size, rank, world)       # this rank's slice of the mixed batch
    got = g.count_mixed(torch.from_numpy(cid[lo:hi]), torch.from_numpy(qs[lo:hi]), torch.from_numpy(qe[lo:hi]), count_fn=count_fn)
    want = np.zeros(hi - lo, np.int64)
    for c in range(len(data)):
        sel = cid[lo:hi] == c
        want[sel] = Oracle(*data[c]).count_batch(qs[lo:hi][sel], qe[lo:hi][sel])
    bases, totals = g.csr_bases()
    full = np.array([int(Oracle(*data[c]).count_batch(qs[cid == c], qe[cid == c]).sum()) for c in range(len(data))])
    ok = np.array_equal(got.numpy().astype(np.int64), want) and set(asked) <= set(g.owned) and np.array_equal(totals, full)
    ok &= g.last_exchange["dispatch_bytes"] > 0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world_size_2_mixed_batch_dispatch_and_combine_gloo():
    """Mode B with a sharded mixed batch: route by destination, all-to-all dispatch, count on the owner, all-to-all
    combine -- the exchange logic on CPU (gloo) with the routing and counting steps injected (numpy / oracle)."""
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_mixed_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs: p.start()
        for p in procs: p.join(180)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}


@pytest.mark.gpu
def test_device_routing_equals_the_host_definition_and_mixed_counts_match():
    import torch
    from oracle.pyoracle import Oracle
    from superintervals_b200.genome import GenomeIndex, _CudaRouter, route_by_contig
    data, cid, qs, qe = _mixed_case(24, 11)
    R = _CudaRouter()
    gs, ge, perm, off = R.route(torch.from_numpy(cid).cuda(), torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda(), 24)
    order, bounds = route_by_contig(cid, 24)
    assert np.array_equal(perm.cpu().numpy().astype(np.int64), order) and np.array_equal(off, bounds)
    assert np.array_equal(gs.cpu().numpy(), qs[order]) and np.array_equal(ge.cpu().numpy(), qe[order])
    g = GenomeIndex([f"c{i}" for i in range(24)], [d[0].size for d in data], rank=0, world=1)
    for c in range(24):
        g.build_contig(c, torch.from_numpy(data[c][0]).cuda(), torch.from_numpy(data[c][1]).cuda())
    got = g.count_mixed(torch.from_numpy(cid).cuda(), torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda())
    want = np.zeros(cid.size, np.int64)
    for c in range(24):
        want[cid == c] = Oracle(*data[c]).count_batch(qs[cid == c], qe[cid == c])
    assert np.array_equal(got.cpu().numpy().astype(np.int64), want)
    with pytest.raises(Exception):
        R.route(torch.from_numpy(np.array([0, 99], np.int32)).cuda(), torch.zeros(2, dtype=torch.int32).cuda(),
                torch.zeros(2, dtype=torch.int32).cuda(), 24)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 3])
def test_per_contig_counts_match_per_contig_oracles(world):
    """C4 scaled down: 24 contigs, queries routed by contig id; the union of what the ranks answer
    equals one oracle per contig (the reference's own per-chromosome convention)."""
    import torch
    from oracle.pyoracle import Oracle
    from superintervals_b200.genome import GenomeIndex
    parts = W.config4_partition(240_000, 600_000)
    names = [f"chr{i + 1}" for i in range(24)]
    data = [W.config4_contig(n_c, q_c, L // 100, seed) for n_c, q_c, L, seed in parts]
    cid = np.concatenate([np.full(d[2].size, c, np.int64) for c, d in enumerate(data)])
    qs = np.concatenate([d[2] for d in data]); qe = np.concatenate([d[3] for d in data])
    perm = np.random.default_rng(1).permutation(cid.size)                 # a mixed, unordered batch
    cid, qs, qe = cid[perm], qs[perm], qe[perm]
    want = np.zeros(cid.size, np.uint64)
    for c, d in enumerate(data):
        sel = cid == c
        want[sel] = Oracle(d[0], d[1]).count_batch(qs[sel], qe[sel])
    got = np.zeros(cid.size, np.uint64)
    totals = np.zeros(24, np.int64)
    for rank in range(world):                                              # ranks simulated one after another on one GPU
        g = GenomeIndex(names, [p[0] for p in parts], [p[1] for p in parts], rank=rank, world=world)
        for c in g.owned:
            g.build_contig(c, torch.from_numpy(data[c][0]).cuda(), torch.from_numpy(data[c][1]).cuda())
        got += g.count(cid, qs, qe)
        totals += g.csr_bases()[1]
    assert np.array_equal(got, want)
    assert np.array_equal(totals, np.array([int(want[cid == c].sum()) for c in range(24)]))


@pytest.mark.gpu
def test_mixed_batch_in_one_launch_matches_per_contig_oracles_incl_malformed_empty_and_inverted():
    """siCountMixedDevice: the whole mixed batch in the caller's order, no routing. Contig 2 stores a few start > end
    intervals (side list), contig 3 has no index, some queries are inverted (qs > qe: the walk's definition) and some
    carry an id outside the table (count 0). Against one oracle per contig."""
    import ctypes as C
    import torch
    from oracle.pyoracle import Oracle
    from superintervals_b200 import _lib
    from superintervals_b200.device import DeviceIndex
    data, cid, qs, qe = _mixed_case(6, 17)
    s2, e2 = data[2]
    s2, e2 = s2.copy(), e2.copy()
    s2[[5, 40, 77]], e2[[5, 40, 77]] = e2[[5, 40, 77]] + 50, s2[[5, 40, 77]]       # start > end
    data[2] = (s2, e2)
    qs, qe = qs.copy(), qe.copy()
    inv = np.arange(0, qs.size, 41)
    qs[inv], qe[inv] = qe[inv] + 3, qs[inv]
    cid = cid.copy()
    cid[7], cid[8] = 99, -1
    idx = [None if c == 3 else DeviceIndex().build(torch.from_numpy(data[c][0]).cuda(), torch.from_numpy(data[c][1]).cuda())
           for c in range(6)]
    L = _lib.lib()
    arr = (C.c_void_p * 6)(*[ix._ix if ix is not None else None for ix in idx])
    out = torch.full((cid.size,), -1, dtype=torch.int32, device="cuda")
    d_cid, d_qs, d_qe = torch.from_numpy(cid).cuda(), torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    tot = torch.full((6,), -1, dtype=torch.int64, device="cuda")
    rc = L.siCountMixedDevice(arr, 6, d_cid.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(), cid.size, out.data_ptr(), tot.data_ptr(), None)
    assert rc == 0
    _lib.check("siCountMixedDevice")
    torch.cuda.synchronize()
    want = np.zeros(cid.size, np.int64)
    for c in range(6):
        if c != 3:
            want[cid == c] = Oracle(*data[c]).count_batch(qs[cid == c], qe[cid == c])
    assert np.array_equal(out.cpu().numpy().astype(np.int64), want)
    assert tot.cpu().numpy().tolist() == [int(want[cid == c].sum()) for c in range(6)]
    # an index that cannot answer from rank cells (walk forced): the call declines without latching an error
    from superintervals_b200.device import OPT_COUNT_ALGO, COUNT_WALK
    idx[0].set_option(OPT_COUNT_ALGO, COUNT_WALK)
    assert L.siCountMixedDevice(arr, 6, d_cid.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(), cid.size, out.data_ptr(), None, None) == -2
    _lib.check("declined mixed count latches nothing")


@pytest.mark.gpu
def test_search_values_of_a_mixed_batch_is_the_contig_major_csr_of_per_contig_oracles():
    import torch
    from oracle.pyoracle import Oracle
    from superintervals_b200.genome import GenomeIndex, route_by_contig
    data, cid, qs, qe = _mixed_case(7, 23)
    g = GenomeIndex([f"c{i}" for i in range(7)], [d[0].size for d in data], rank=0, world=1)
    for c in range(7):
        if c != 4:                                        # contig 4 has no index here: empty lists
            g.build_contig(c, torch.from_numpy(data[c][0]).cuda(), torch.from_numpy(data[c][1]).cuda())
    perm, off, vals = g.search_values_mixed(torch.from_numpy(cid).cuda(), torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda())
    order, bounds = route_by_contig(cid, 7)
    assert np.array_equal(perm.cpu().numpy().astype(np.int64), order)
    off, vals = off.cpu().numpy(), vals.cpu().numpy()
    want_vals, want_counts = [], []
    for c in range(7):
        sel = order[bounds[c]:bounds[c + 1]]
        if c == 4 or sel.size == 0:
            want_counts.append(np.zeros(sel.size, np.int64))
            continue
        o, res = Oracle(*data[c]).search_batch(qs[sel], qe[sel], want=("values",))
        want_counts.append(np.diff(o.astype(np.int64)))
        want_vals.append(res["values"])
    want_off = np.concatenate([[0], np.cumsum(np.concatenate(want_counts))])
    assert np.array_equal(off, want_off)
    assert np.array_equal(vals, np.concatenate(want_vals))
    bases, totals = g.csr_bases()
    assert np.array_equal(totals, [int(c.sum()) for c in want_counts]) and np.array_equal(bases, want_off[bounds[:-1]])
