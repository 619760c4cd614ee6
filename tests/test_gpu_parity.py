"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Bit-exact bar: counts, CSR offsets, value/idx/key lists (including their
descending order), the index arrays produced by build(), has_overlaps quirk Q1.
Seeded inputs at sizes the oracle finishes in seconds.
"""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from superintervals_b200 import workloads as W
from superintervals_b200 import _lib

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
    """All count kernels (branch-array walk / closed-form rank by grid / by rank cells), every partition shape (0-3 passes,
    with and without result windows) and every order mode give the oracle's counts; the CSR fill
    reuses each partition. Inverted queries inside a rank batch take the walk (quirk Q6)."""
    import torch
    from superintervals_b200.device import (COUNT_CELLS, COUNT_RANK, COUNT_WALK, OPT_BUCKET_INTERVALS,
                                            OPT_CELLS_DIRECT_BYTES, OPT_COUNT_ALGO, OPT_WINDOW_SHIFT, ORDER_ASIS,
                                            ORDER_SORTED, ORDER_UNSORTED, DeviceIndex)
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
    for algo, direct in ((COUNT_WALK, 0), (COUNT_RANK, 0), (COUNT_CELLS, 0), (COUNT_CELLS, 1)):
        # cells: once answering the unpartitioned batch straight from L2, once through the partition
        ix.set_option(OPT_COUNT_ALGO, algo).set_option(OPT_CELLS_DIRECT_BYTES, direct)
        assert np.array_equal(u32(ix.count(dqs, dqe, order=ORDER_ASIS)), want), (algo, "asis")
        assert np.array_equal(u32(ix.count(sqs, sqe, order=ORDER_SORTED)), want[srt]), (algo, "sorted")
        for bucket, wshift in PLANS:
            ix.set_option(OPT_BUCKET_INTERVALS, bucket).set_option(OPT_WINDOW_SHIFT, wshift)
            assert np.array_equal(u32(ix.count(dqs, dqe, order=ORDER_UNSORTED)), want), (algo, bucket, wshift)
            off, vals = ix.search_values(dqs, dqe, order=ORDER_UNSORTED)
            assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o), (algo, bucket, wshift)
            assert np.array_equal(vals.cpu().numpy(), res["values"]), (algo, bucket, wshift)


def _brute_counts(s, e, qs, qe):
    """#{ j : s[j] <= qe, e[j] >= qs } for well-formed intervals and qs <= qe (closed form, int64)."""
    ss, es = np.sort(s.astype(np.int64)), np.sort(e.astype(np.int64))
    return (np.searchsorted(ss, qe.astype(np.int64), "right") - np.searchsorted(es, qs.astype(np.int64), "left")).astype(np.uint64)


CELL_CASES = ["full_range", "one_point", "clusters", "dense", "sparse", "single", "two_far"]


@pytest.mark.parametrize("case", CELL_CASES)
@pytest.mark.parametrize("fill", [0, 1, 3, 28])
def test_rank_cells_edge_cases(case, fill):
    """The rank cells (one 32-byte record per 2^k coordinates) at their edges: the whole int32 span,
    every interval on one coordinate (over-full cells fall back to the sorted arrays), clustered starts,
    both record formats at several cell widths, queries at INT_MIN / INT_MAX and inverted queries."""
    import torch
    from superintervals_b200.device import (COUNT_CELLS, COUNT_WALK, OPT_CELLS_FILL, OPT_COUNT_ALGO, ORDER_ASIS,
                                            DeviceIndex)
    rng = np.random.default_rng(100 * CELL_CASES.index(case) + fill)
    I32 = np.iinfo(np.int32)
    n, nq = 40_000, 60_000
    if case == "full_range":
        s = rng.integers(I32.min, I32.max, n, dtype=np.int64)
        e = np.minimum(s + rng.integers(0, 1 << 28, n), I32.max)
        s[0], e[0] = I32.min, I32.min
        s[1], e[1] = I32.max, I32.max
        s[2], e[2] = I32.min, I32.max
    elif case == "one_point":
        s = np.full(n, 12345, np.int64); e = s.copy()
    elif case == "clusters":
        c = rng.integers(0, 20, n) * 1_000_000
        s = c + rng.integers(0, 40, n); e = s + rng.integers(0, 300, n)
    elif case == "dense":
        s = rng.integers(0, 4_000, n, dtype=np.int64); e = s + rng.integers(0, 50, n)
    elif case == "sparse":
        s = rng.integers(-1_000_000_000, 1_000_000_000, 3_000, dtype=np.int64); e = s + rng.integers(0, 5_000_000, 3_000)
    elif case == "single":
        s = np.array([7], np.int64); e = np.array([9], np.int64)
    else:
        s = np.array([I32.min, I32.max - 5], np.int64); e = np.array([I32.min + 3, I32.max], np.int64)
    s, e = s.astype(np.int32), e.astype(np.int32)
    lo, hi = int(s.min()), int(e.max())
    near = np.clip(rng.integers(lo - 300, hi + 300, nq - 8, dtype=np.int64), I32.min, I32.max)
    pick = rng.integers(0, s.size, nq - 8)
    mix = rng.random(nq - 8)
    qs = np.where(mix < 0.5, near, np.where(mix < 0.75, s[pick], e[pick]).astype(np.int64) + rng.integers(-2, 3, nq - 8))
    qs = np.clip(qs, I32.min, I32.max)
    qe = np.clip(qs + rng.integers(0, 1 << int(rng.integers(1, 30)), nq - 8), I32.min, I32.max)
    qs = np.concatenate([qs, [I32.min, I32.min, I32.max, I32.max, lo, hi, hi, I32.min]])
    qe = np.concatenate([qe, [I32.min, I32.max, I32.max, I32.max, lo, hi, I32.max, lo]])
    inv = rng.random(qs.size) < 0.02                       # a few inverted queries (quirk Q6) take the walk
    qs, qe = np.where(inv, qe, qs).astype(np.int32), np.where(inv, qs, qe).astype(np.int32)

    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    u32 = lambda t: t.cpu().numpy().astype(np.uint32).astype(np.uint64)
    ix = DeviceIndex()
    if fill:
        ix.set_option(OPT_CELLS_FILL, fill)
    ix.build(dev(s), dev(e))
    info = ix.cells_info()
    assert info["starts"]["format"] in (1, 2) and info["ends"]["format"] in (1, 2)
    if case == "one_point":
        assert info["starts"]["overfull"] >= 1
    dqs, dqe = dev(qs), dev(qe)
    got = u32(ix.set_option(OPT_COUNT_ALGO, COUNT_CELLS).count(dqs, dqe, order=ORDER_ASIS))
    walk = u32(ix.set_option(OPT_COUNT_ALGO, COUNT_WALK).count(dqs, dqe, order=ORDER_ASIS))
    assert np.array_equal(got, walk)
    ok = qs <= qe
    assert np.array_equal(got[ok], _brute_counts(s, e, qs[ok], qe[ok]))
    orc = Oracle(s, e)
    want = orc.count_batch(qs, qe)
    assert np.array_equal(got, want)
    # CSR fill by run + stab (qk_fill_runs_kernel) on the same edges, every payload, against the oracle and
    # against the plain walk fill; a prefix of the batch whose lists stay small enough to compare
    from superintervals_b200.device import FILL_IDXS, FILL_ITEMS, FILL_KEYS, FILL_VALUES
    m = max(1, int(np.searchsorted(np.cumsum(want), 3_000_000)))
    off_o, res = orc.search_batch(qs[:m], qe[:m], want=("values", "idxs", "keys"))
    pqs, pqe = dev(qs[:m]), dev(qe[:m])
    from superintervals_b200.device import OPT_STAB_BUDGET, OPT_STAB_LISTS
    for what, key in ((FILL_VALUES, "values"), (FILL_IDXS, "idxs"), (FILL_KEYS, "keys"), (FILL_ITEMS, None)):
        off_w, lst_w = ix.set_option(OPT_COUNT_ALGO, COUNT_WALK).search(pqs, pqe, what, order=ORDER_ASIS)
        ix.set_option(OPT_COUNT_ALGO, COUNT_CELLS)
        # below each run: the stab lists, the branch-array walk, and lists at a coarser checkpoint spacing
        # ... and the (value, end) copy of the short lists on and off (a view switch: the lists themselves stay)
        for lists, budget, vlists in ((1, 6, 1), (1, 6, 0), (0, 6, 1), (1, 1, 1)):
            ix.set_option(OPT_STAB_LISTS, lists).set_option(OPT_STAB_BUDGET, budget).set_option(_lib.OPT_STAB_VALUE_LISTS, vlists)
            off, lst = ix.search(pqs, pqe, what, order=ORDER_ASIS)
            assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o), (case, what, lists, budget, vlists)
            assert torch.equal(off, off_w) and torch.equal(lst, lst_w), (case, what, lists, budget, vlists)
        ix.set_option(_lib.OPT_STAB_VALUE_LISTS, 1)
        if key:
            assert np.array_equal(lst_w.cpu().numpy().astype(res[key].dtype).reshape(res[key].shape), res[key]), (case, key)
    ix.set_option(OPT_STAB_LISTS, 1).set_option(OPT_STAB_BUDGET, 6)


@pytest.mark.parametrize("shape,budget,want_state,min_shift", [("c2", 6, 1, 4), ("c2", 1, 1, 6), ("c3", 6, 1, 3),
                                                               ("deep", 6, 2, 0), ("deep", 4096, 1, 3), ("tiny", 6, 2, 0)])
def test_stab_lists_budget_and_fallback(shape, budget, want_state, min_shift):
    """The stab lists behind the CSR fill: built on the first fill, checkpoint spacing doubled until the
    lists fit the budget, an index nested too deeply (or too small) keeps the branch-array walk; every
    variant returns the oracle's lists, and a rebuild drops the old lists."""
    import torch
    from superintervals_b200.device import (FILL_KEYS, OPT_STAB_BUDGET, ORDER_ASIS, DeviceIndex)
    rng = np.random.default_rng(5)
    if shape == "c2":
        s, e, qs, qe = W.config2(60_000, 40_000, 3, axis=1_500_000)          # ~100 open intervals at every point
    elif shape == "c3":
        s, e, qs, qe = W.config3(80_000, 40_000, 9, axis=5_000_000)
    elif shape == "deep":                                                    # every interval contains the next: depth = n
        n = 40_000
        s = np.arange(n, dtype=np.int32); e = (2 * n - s).astype(np.int32)
        qs = rng.integers(-5, 2 * n + 5, 300).astype(np.int32); qe = (qs + rng.integers(0, 50, 300)).astype(np.int32)
    else:
        s = np.arange(40, dtype=np.int32); e = (s + 3).astype(np.int32)
        qs = rng.integers(-2, 50, 500).astype(np.int32); qe = (qs + rng.integers(0, 9, 500)).astype(np.int32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ix = DeviceIndex().set_option(OPT_STAB_BUDGET, budget).build(dev(s), dev(e))
    assert ix.stab_info()["state"] == 0
    o = Oracle(s, e)
    off_o, res = o.search_batch(qs, qe, want=("values", "keys"))
    off, vals = ix.search_values(dev(qs), dev(qe), order=ORDER_ASIS)
    info = ix.stab_info()
    assert info["state"] == want_state, info
    if want_state == 1:
        assert info["shift"] >= min_shift and info["entries"] <= budget * s.size and info["lists"] == (s.size >> info["shift"]) + 1
    assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o)
    assert np.array_equal(vals.cpu().numpy(), res["values"])
    _, keys = ix.search(dev(qs), dev(qe), FILL_KEYS, order=ORDER_ASIS)
    assert np.array_equal(keys.cpu().numpy(), res["keys"])
    # a rebuild forgets the lists; the next fill makes them for the new index
    s2, e2 = (s[: s.size // 2] + 7).astype(np.int32), (e[: s.size // 2] + 9).astype(np.int32)
    ix.build(dev(s2), dev(e2))
    assert ix.stab_info()["state"] == 0
    off2, vals2 = ix.search_values(dev(qs), dev(qe), order=ORDER_ASIS)
    off_o2, res2 = Oracle(s2, e2).search_batch(qs, qe)
    assert np.array_equal(off2.cpu().numpy().astype(np.uint64), off_o2) and np.array_equal(vals2.cpu().numpy(), res2["values"])
