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
