"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Bit-exact bar: counts, CSR offsets, value/idx/key lists (including their
descending order), the index arrays produced by build(), has_overlaps quirk Q1.
Seeded inputs at sizes the oracle finishes in seconds.
"""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from superintervals_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _imap(starts, ends):
    from superintervals_b200 import IntervalMap
    return IntervalMap.from_arrays(starts, ends)


def _mk(kind, n, nq, seed):
    rng = np.random.default_rng(seed)
    if kind == "c1":
        return W.config1(n, seed)
    if kind == "c2":
        return W.config2(n, nq, seed, axis=max(20_000, 25 * n))
    if kind == "c3":
        return W.config3(n, nq, seed, axis=max(2_000_000, 60 * n))
    if kind == "dups":     # many exact duplicates and shared endpoints
        s = rng.integers(0, 50, n).astype(np.int32)
        e = (s + rng.integers(0, 8, n)).astype(np.int32)
        qs = rng.integers(-5, 60, nq).astype(np.int32)
        qe = (qs + rng.integers(0, 12, nq)).astype(np.int32)
        return s, e, qs, qe
    if kind == "nested":   # one giant container + a rising staircase under it
        s = np.arange(n, dtype=np.int32)
        e = (s + 1 + (np.arange(n) % 7)).astype(np.int32)
        s[0], e[0] = 0, 10 * n
        qs = rng.integers(0, n + 10, nq).astype(np.int32)
        qe = (qs + rng.integers(0, 40, nq)).astype(np.int32)
        return s, e, qs, qe
    if kind == "negative":  # signed coordinates, inverted queries (Q6/Q7)
        s = rng.integers(-2_000_000_000, 2_000_000_000, n).astype(np.int64)
        e = np.minimum(s + rng.integers(0, 50_000_000, n), 2_147_483_000).astype(np.int32)
        s = s.astype(np.int32)
        qs = rng.integers(-2_100_000_000, 2_100_000_000, nq).astype(np.int64)
        qe = np.clip(qs + rng.integers(-1000, 80_000_000, nq), -2_147_483_648, 2_147_483_647).astype(np.int32)
        return s, e, qs.astype(np.int32), qe
    raise KeyError(kind)


CASES = [("c1", 20_000, 20_000, 0), ("c1", 200_000, 200_000, 1), ("c2", 50_000, 300_000, 2),
         ("c3", 100_000, 100_000, 42), ("dups", 5_000, 8_000, 7), ("nested", 70_000, 30_000, 9),
         ("negative", 30_000, 30_000, 11), ("c1", 33, 100, 3), ("c1", 1, 10, 4), ("dups", 129, 500, 5)]


@pytest.mark.parametrize("kind,n,nq,seed", CASES)
@pytest.mark.parametrize("presorted", [False, True])
def test_build_matches_oracle(kind, n, nq, seed, presorted):
    s, e, _, _ = _mk(kind, n, nq, seed)
    if presorted:
        s, e = W.sort_by_start(s, e)
    o = Oracle(s, e)
    m = _imap(s, e)
    assert np.array_equal(m.starts, o.starts)
    assert np.array_equal(m.ends, o.ends)
    assert np.array_equal(m.data_index, o.data)      # stable tie order (Q3)
    assert np.array_equal(m.branch, o.branch)


@pytest.mark.parametrize("kind,n,nq,seed", CASES)
@pytest.mark.parametrize("sort_queries", [False, True])
def test_count_and_search_match_oracle(kind, n, nq, seed, sort_queries):
    s, e, qs, qe = _mk(kind, n, nq, seed)
    if sort_queries:
        order = np.argsort(qs, kind="stable")
        qs, qe = qs[order], qe[order]
    o = Oracle(s, e)
    m = _imap(s, e)
    want = o.count_batch(qs, qe)
    got = m.count_batch_np(qs, qe)
    assert np.array_equal(got, want)
    assert np.array_equal(m.has_overlaps_batch(qs, qe), o.has_overlaps_batch(qs, qe))
    off_o, res = o.search_batch(qs, qe, want=("values", "idxs", "keys"))
    off, vals = m.search_values_batch_csr(qs, qe)
    assert np.array_equal(off, off_o)
    assert np.array_equal(vals, res["values"])
    off2, idx = m.search_idxs_batch_csr(qs, qe)
    assert np.array_equal(off2, off_o) and np.array_equal(idx.astype(np.uint32), res["idxs"])
    off3, keys = m.search_keys_batch_csr(qs, qe)
    assert np.array_equal(off3, off_o) and np.array_equal(keys, res["keys"])


PLANS = [(1024, 22), (1, 10), (64, 12), (1 << 20, 22), (16, 16)]   # (bucket_intervals, window_shift): 0..3 partition passes


@pytest.mark.parametrize("kind,n,nq,seed", [("c1", 200_000, 200_000, 1), ("c2", 50_000, 300_000, 2),
                                            ("c3", 100_000, 100_000, 42), ("dups", 5_000, 8_000, 7),
                                            ("nested", 70_000, 30_000, 9), ("negative", 30_000, 30_000, 11),
                                            ("c1", 1, 10, 4)])
def test_count_kernels_and_partition_plans_agree(kind, n, nq, seed):
    """Both count kernels (branch-array walk / closed-form rank), every partition shape (0-3 passes,
    with and without result windows) and every order mode give the oracle's counts; the CSR fill
    reuses each partition. Inverted queries inside a rank batch take the walk (quirk Q6)."""
    import torch
    from superintervals_b200.device import (COUNT_RANK, COUNT_WALK, OPT_BUCKET_INTERVALS, OPT_COUNT_ALGO,
                                            OPT_WINDOW_SHIFT, ORDER_ASIS, ORDER_SORTED, ORDER_UNSORTED, DeviceIndex)
    s, e, qs, qe = _mk(kind, n, nq, seed)
    o = Oracle(s, e)
    want = o.count_batch(qs, qe)
    off_o, res = o.search_batch(qs, qe)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    u32 = lambda t: t.cpu().numpy().astype(np.uint32).astype(np.uint64)
    ix = DeviceIndex().build(dev(s), dev(e))
    dqs, dqe = dev(qs), dev(qe)
    srt = np.argsort(qs, kind="stable")
    sqs, sqe = dev(qs[srt]), dev(qe[srt])
    for algo in (COUNT_WALK, COUNT_RANK):
        ix.set_option(OPT_COUNT_ALGO, algo)
        assert np.array_equal(u32(ix.count(dqs, dqe, order=ORDER_ASIS)), want), (algo, "asis")
        assert np.array_equal(u32(ix.count(sqs, sqe, order=ORDER_SORTED)), want[srt]), (algo, "sorted")
        for bucket, wshift in PLANS:
            ix.set_option(OPT_BUCKET_INTERVALS, bucket).set_option(OPT_WINDOW_SHIFT, wshift)
            assert np.array_equal(u32(ix.count(dqs, dqe, order=ORDER_UNSORTED)), want), (algo, bucket, wshift)
            off, vals = ix.search_values(dqs, dqe, order=ORDER_UNSORTED)
            assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o), (algo, bucket, wshift)
            assert np.array_equal(vals.cpu().numpy(), res["values"]), (algo, bucket, wshift)
