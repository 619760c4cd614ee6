"""GPU: the multi-device entry points of the C ABI (csrc/multi.cu) on every visible device -- one device on
the round-end box, several under `gpurun --gpus N`. Same answers as the oracle; NCCL traffic whenever more
than one device takes part; the C example drives the same entry points from plain C."""
import os
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import Oracle
from superintervals_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "multi_count")


def _compile():
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", os.path.join(ROOT, "examples", "multi_count.c"), "-I" + os.path.join(ROOT, "include"),
           "-L" + os.path.join(ROOT, "superintervals_b200"), "-lsuperintervals_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "superintervals_b200"), "-o", EXE]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_multi_count_and_search_match_the_oracle_on_every_visible_device():
    import torch
    from superintervals_b200.multi import MultiIndex
    s, e, qs, qe = W.config3(200_000, 150_001, 42, axis=12_000_000)
    orc = Oracle(s, e)
    want = orc.count_batch(qs, qe)
    off_o, res = orc.search_batch(qs, qe, want=("values",))
    for n in sorted({1, torch.cuda.device_count()}):
        m = MultiIndex(n).build(s, e)
        assert m.n_devices == n
        got = m.count_batch(qs, qe)
        assert np.array_equal(got.astype(np.uint64), want), n
        st = m.stats()
        assert (st["nccl_bytes"] > 0) == (n > 1)
        off, vals = m.search_values_batch_csr(qs, qe)
        assert np.array_equal(off, off_o) and np.array_equal(vals, res["values"]), n
        # tiny and empty batches, and a batch shorter than the device count
        for k in (0, 1, 3, 17):
            assert np.array_equal(m.count_batch(qs[:k], qe[:k]).astype(np.uint64), want[:k])


def test_c_program_drives_the_multi_device_entry_points():
    _compile()
    out = subprocess.run([EXE, "0", "200000", "500003"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout and "identical offsets and values" in out.stdout and "multi ok" in out.stdout


def _fused_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from superintervals_b200 import workloads as W
    from superintervals_b200.device import DeviceIndex, ORDER_UNSORTED
    from superintervals_b200.sharding import PeerGathered
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s, e, qs, qe = W.config2(50_000, 200_003, 5, axis=2_000_000)
    ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    nq = qs.size
    per = (((nq + world - 1) // world) + 7) & ~7
    lo = min(nq, rank * per); hi = min(nq, lo + per); m = hi - lo
    mqs, mqe = torch.from_numpy(qs[lo:hi]).cuda(), torch.from_numpy(qe[lo:hi]).cuda()
    pg = PeerGathered(per, world, rank)
    ok = True
    want = Oracle(s, e).count_batch(qs, qe).astype(np.int64)
    for step in range(4):                      # both halves of the double buffer, twice
        arr, slot, ptrs = pg.current()
        ix.count_fanout(mqs, mqe, slot[:m], ptrs, order=ORDER_UNSORTED)
        pg.barrier()
        torch.cuda.synchronize()
        got = torch.cat([arr[r * per: r * per + max(0, min(nq, (r + 1) * per) - r * per)] for r in range(world)]).cpu().numpy()
        ok &= bool(np.array_equal(got.astype(np.int64), want))
    ok &= not pg.timed_out()
    pg.close()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_fused_count_and_all_gather_over_peer_memory_two_processes():
    """One process per GPU (torchrun's layout): each rank's count kernel stores its slot of the gathered count vector into
    every GPU's copy (CUDA IPC peer memory over NVLink), a one-warp kernel exchanges flags; every rank then holds the
    whole batch's counts -- equal to the oracle's. Needs two GPUs."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs: p.start()
        for p in procs: p.join(180)
        for p in procs:
            if p.is_alive(): p.kill()       # a rank that died leaves its peer in a collective: do not wait for it
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}


def _peer_mixed_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from superintervals_b200 import workloads as W
    from superintervals_b200.genome import GenomeIndex
    from superintervals_b200.sharding import PeerBatch
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    axes = [3_000_000, 1_500_000, 800_000, 2_000_000, 500_000]
    ns = [60_000, 40_000, 0, 50_000, 20_000]                      # contig 2 has no intervals: nobody indexes it
    g = GenomeIndex([f"c{i}" for i in range(5)], ns, rank=rank, world=world, pair_cells=2 if rank == 0 else 0)
    orcs = {}
    for c, (n, axis) in enumerate(zip(ns, axes)):
        if n == 0:
            continue
        s, e = W.config2_intervals(n, 30 + c, axis=axis)
        orcs[c] = Oracle(s, e)
        if g.owns(c):
            g.build_contig(c, torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    ok = len(g.owned) > 0
    per = 150_000 + 1000 * rank                                    # slices of different lengths
    batch = PeerBatch(150_000 + 1000 * (world - 1), world, rank)   # one capacity for every rank: the longest slice
    for step in range(3):
        rng = np.random.default_rng([step, rank])
        cid = rng.integers(0, 6, per).astype(np.int32)             # id 5 lies outside the table
        qs, qe = W.config2_queries(per, 40 + step, axis=3_000_000, shard=rank)
        inv = rng.random(per) < 0.001                              # a few inverted queries (quirk Q6)
        qs, qe = np.where(inv, qe, qs).astype(np.int32), np.where(inv, qs, qe).astype(np.int32)
        batch.contig[:per].copy_(torch.from_numpy(cid)); batch.qs[:per].copy_(torch.from_numpy(qs)); batch.qe[:per].copy_(torch.from_numpy(qe))
        batch.counts.fill_(-1)
        batch.set_length(per)
        # step 0: persistent CTAs, a thread per query (tables in L2); steps 1, 2: short-lived CTAs of 8 / 3 tiles that pack their
        # own queries into a shared list first (what a whole genome's tables get)
        os.environ.pop("SIB_QM_ROUNDS", None)
        if step:
            os.environ["SIB_QM_ROUNDS"] = "8" if step == 1 else "3"
        got = g.count_mixed_peer(batch)
        ok &= got is not None
        if got is None:
            break
        want = np.zeros(per, np.int64)
        for c, orc in orcs.items():
            m = cid == c
            want[m] = orc.count_batch(qs[m], qe[m]).astype(np.int64)
        ok &= bool(np.array_equal(got.cpu().numpy().astype(np.uint32).astype(np.int64), want))
        # the owned contigs' totals cover the WHOLE batch: compare their sum over the ranks with the batch's hits
        tot = torch.tensor([int(g.hits.sum()), int(want.sum())], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        ok &= int(tot[0]) == int(tot[1])
    ok &= not batch.timed_out()
    batch.close()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_contig_partitioned_mixed_batch_read_in_place_over_peer_memory_two_processes():
    """Mode B without a dispatch (siCountMixedPeerDevice): contigs partitioned over two GPUs, each rank's slice of the
    mixed batch stays in its own (IPC-shared) memory, every rank answers the queries of its contigs in place and stores
    the counts into the slice they belong to. Equal to one oracle per contig, empty / unknown contigs count 0. Needs two GPUs."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_peer_mixed_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs: p.start()
        for p in procs: p.join(180)
        for p in procs:
            if p.is_alive(): p.kill()
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}
