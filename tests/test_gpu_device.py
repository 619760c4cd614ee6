"""GPU: the device-resident core (raw pointers + stream) -- order modes, the explicit
sort/count split, CSR scan, and full-size runs checked through size-independent properties."""
import numpy as np
import pytest
import torch

from oracle.pyoracle import Oracle
from superintervals_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _u32(t):
    return t.cpu().numpy().astype(np.uint32).astype(np.uint64)


@pytest.fixture(scope="module")
def c3():
    from superintervals_b200.device import DeviceIndex
    s, e, qs, qe = W.config3(300_000, 200_000, 42, axis=20_000_000)
    return s, e, qs, qe, DeviceIndex().build(_dev(s), _dev(e)), Oracle(s, e)


def test_every_order_mode_gives_the_callers_order(c3):
    from superintervals_b200.device import ORDER_ASIS, ORDER_AUTO, ORDER_SORTED, ORDER_UNSORTED
    s, e, qs, qe, ix, orc = c3
    want = orc.count_batch(qs, qe)
    dqs, dqe = _dev(qs), _dev(qe)
    for order in (ORDER_AUTO, ORDER_UNSORTED, ORDER_ASIS):
        assert np.array_equal(_u32(ix.count(dqs, dqe, order=order)), want), order
    ix.sort_queries(dqs, dqe)                                  # explicit two-phase form
    assert np.array_equal(_u32(ix.count(dqs, dqe, order=ORDER_UNSORTED)), want)
    o = np.argsort(qs, kind="stable")
    assert np.array_equal(_u32(ix.count(_dev(qs[o]), _dev(qe[o]), order=ORDER_SORTED)), want[o])


def test_unsorted_count_resorts_when_the_batch_changes_in_place(c3):
    from superintervals_b200.device import ORDER_UNSORTED
    s, e, qs, qe, ix, orc = c3
    dqs, dqe = _dev(qs[:50_000]), _dev(qe[:50_000])
    assert np.array_equal(_u32(ix.count(dqs, dqe, order=ORDER_UNSORTED)), orc.count_batch(qs[:50_000], qe[:50_000]))
    dqs.copy_(_dev(qs[50_000:100_000])); dqe.copy_(_dev(qe[50_000:100_000]))   # same pointers, new contents
    assert np.array_equal(_u32(ix.count(dqs, dqe, order=ORDER_UNSORTED)),
                          orc.count_batch(qs[50_000:100_000], qe[50_000:100_000]))


def test_search_csr_on_device(c3):
    from superintervals_b200._lib import FILL_ITEMS
    s, e, qs, qe, ix, orc = c3
    off_o, res = orc.search_batch(qs, qe, want=("values", "idxs", "keys"))
    dqs, dqe = _dev(qs), _dev(qe)
    off, vals = ix.search_values(dqs, dqe)
    assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o)
    assert np.array_equal(vals.cpu().numpy(), res["values"])
    _, idx = ix.search_idxs(dqs, dqe)
    assert np.array_equal(idx.cpu().numpy().astype(np.uint32), res["idxs"])
    _, keys = ix.search_keys(dqs, dqe)
    assert np.array_equal(keys.cpu().numpy(), res["keys"])
    _, items = ix.search(dqs, dqe, FILL_ITEMS)
    assert np.array_equal(items.cpu().numpy()[:, :2], res["keys"]) and np.array_equal(items.cpu().numpy()[:, 2], res["values"])
    cnt, cov = ix.coverage(dqs, dqe)
    assert np.array_equal(_u32(cnt), orc.count_batch(qs, qe))
    assert np.array_equal(ix.has_overlaps(dqs, dqe).cpu().numpy(), orc.has_overlaps_batch(qs, qe))


def test_scan_matches_cumsum_at_awkward_sizes():
    from superintervals_b200.device import DeviceIndex
    ix = DeviceIndex()
    g = torch.Generator(device="cuda").manual_seed(1)
    for n in (1, 15, 16, 17, 4095, 4096, 4097, 1_000_003):
        c = torch.randint(0, 1 << 20, (n,), device="cuda", dtype=torch.int32, generator=g)
        off = ix.scan(c)
        want = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        want[1:] = torch.cumsum(c.to(torch.int64), 0)
        assert torch.equal(off, want), n


def test_malformed_intervals_disable_the_shortcut_but_stay_exact():
    """start > end is accepted by the reference (Q6); the walk definition still holds verbatim."""
    from superintervals_b200.device import DeviceIndex
    rng = np.random.default_rng(5)
    n, nq = 40_000, 60_000
    s = rng.integers(0, 1_000_000, n).astype(np.int32)
    e = (s + rng.integers(-300, 3000, n)).astype(np.int32)
    qs = rng.integers(0, 1_000_000, nq).astype(np.int32)
    qe = (qs + rng.integers(-50, 4000, nq)).astype(np.int32)
    orc = Oracle(s, e)
    ix = DeviceIndex().build(_dev(s), _dev(e))
    assert np.array_equal(_u32(ix.count(_dev(qs), _dev(qe))), orc.count_batch(qs, qe))
    off_o, res = orc.search_batch(qs, qe)
    off, vals = ix.search_values(_dev(qs), _dev(qe))
    assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o) and np.array_equal(vals.cpu().numpy(), res["values"])


@pytest.mark.parametrize("n_mal", [1, 8, 9])
def test_a_few_malformed_intervals_keep_the_closed_form(n_mal):
    """one start > end interval in 1 M must not flip the index to the walk kernels. Up to 8 are
    listed beside the rank tables (count stays on the cells kernel); more than 8 fall back to the walk. Bit-exact either way,
    inverted queries included (hpp:651-658, quirk Q6)."""
    from superintervals_b200.device import DeviceIndex, OPT_TIMING, ORDER_UNSORTED
    rng = np.random.default_rng(77 + n_mal)
    n, nq = 1_000_000, 300_000
    s = rng.integers(0, 50_000_000, n).astype(np.int32)
    e = (s + rng.integers(0, 3000, n)).astype(np.int32)
    bad = rng.choice(n, n_mal, replace=False)
    e[bad] = s[bad] - rng.integers(1, 5000, n_mal).astype(np.int32)        # start > end
    qs = rng.integers(0, 50_000_000, nq).astype(np.int32)
    qe = (qs + rng.integers(-20, 4000, nq)).astype(np.int32)                # some queries inverted too
    # queries that bracket the malformed intervals from every side
    qs[:n_mal], qe[:n_mal] = e[bad], s[bad]
    qs[n_mal:2 * n_mal], qe[n_mal:2 * n_mal] = s[bad], s[bad]
    qs[2 * n_mal:3 * n_mal], qe[2 * n_mal:3 * n_mal] = e[bad] - 10, e[bad]
    orc = Oracle(s, e)
    ix = DeviceIndex().build(_dev(s), _dev(e))
    ix.set_option(OPT_TIMING, 1)
    got = _u32(ix.count(_dev(qs), _dev(qe), order=ORDER_UNSORTED))
    names = {k for k, _ in ix.read_timings()}
    ix.set_option(OPT_TIMING, 0)
    assert np.array_equal(got, orc.count_batch(qs, qe))
    assert ("count_cells" in names) == (n_mal <= 8), names
    m = 50_000
    off_o, res = orc.search_batch(qs[:m], qe[:m])
    off, vals = ix.search_values(_dev(qs[:m]), _dev(qe[:m]))
    assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o) and np.array_equal(vals.cpu().numpy(), res["values"])
    s1, e1, _, b1, _ = ix.export()
    assert np.array_equal(s1, orc.starts) and np.array_equal(e1, orc.ends) and np.array_equal(b1, orc.branch)


def test_build_export_roundtrip_and_idempotence():
    from superintervals_b200.device import DeviceIndex
    s, e, _, _ = W.config2(200_000, 10, 3, axis=5_000_000)
    ix = DeviceIndex().build(_dev(s), _dev(e))
    s1, e1, v1, b1, p1 = ix.export()
    assert np.array_equal(s1, s[p1]) and np.array_equal(e1, e[p1]) and np.array_equal(v1.astype(np.uint32), p1)
    k = s1.astype(np.int64) * (1 << 32) - e1.astype(np.int64)
    assert (np.diff(k) >= 0).all()                               # (start asc, end desc)
    ix2 = DeviceIndex().build(_dev(s1), _dev(e1))                # rebuilding a built index changes nothing
    s2, e2, _, b2, p2 = ix2.export()
    assert np.array_equal(s2, s1) and np.array_equal(e2, e1) and np.array_equal(b2, b1)
    assert np.array_equal(p2, np.arange(len(s), dtype=np.uint32))
    assert np.array_equal(b1, Oracle(s, e).branch)


@pytest.mark.parametrize("nq", [20_000_000])
def test_full_size_c2_properties(nq):
    """BASELINE configs[1] at full index size: 10M intervals. Checked without the oracle's
    per-query walk: closed-form rank count on GPU, invariance under query order, oracle on a sample."""
    from superintervals_b200.device import DeviceIndex, ORDER_SORTED, ORDER_UNSORTED
    s, e = W.config2_intervals(10_000_000, 2)
    qs, qe = W.config2_queries(nq, 2)
    ds, de, dqs, dqe = _dev(s), _dev(e), _dev(qs), _dev(qe)
    ix = DeviceIndex().build(ds, de)
    c = ix.count(dqs, dqe, order=ORDER_UNSORTED)
    # closed form (well-formed data): #{starts <= qe} - #{ends < qs}
    ss, se = torch.sort(ds)[0], torch.sort(de)[0]
    want = torch.searchsorted(ss, dqe, right=True) - torch.searchsorted(se, dqs, right=False)
    assert torch.equal(c.to(torch.int64), want.to(torch.int64))
    # sorted order gives the same multiset, position by position after permuting
    o = torch.argsort(dqs, stable=True)
    c2 = ix.count(dqs[o].contiguous(), dqe[o].contiguous(), order=ORDER_SORTED)
    assert torch.equal(c2, c[o])
    # oracle on a sample + index structure on the whole build
    orc = Oracle(s, e)
    m = 100_000
    assert np.array_equal(_u32(c[:m]), orc.count_batch(qs[:m], qe[:m]))
    s1, e1, _, b1, _ = ix.export()
    assert np.array_equal(s1, orc.starts) and np.array_equal(e1, orc.ends) and np.array_equal(b1, orc.branch)
    # search_values on a slice: offsets are the scan of counts, values resolve to overlapping intervals
    off, vals = ix.search_values(dqs[:200_000].contiguous(), dqe[:200_000].contiguous())
    assert torch.equal(off[1:] - off[:-1], c[:200_000].to(torch.int64))
    seg = torch.repeat_interleave(torch.arange(200_000, device="cuda"), c[:200_000].to(torch.int64))
    v = vals.to(torch.int64)
    assert bool(((ds[v] <= dqe[seg]) & (de[v] >= dqs[seg])).all())


def test_host_batch_pipeline_large_n():
    """countOverlapsBatch switches to the chunked 3-stream pipeline above 12M queries: same answers."""
    from superintervals_b200 import IntervalMap
    s, e = W.config2_intervals(200_000, 3, axis=5_000_000)
    qs, qe = W.config2_queries(20_000_000, 3, axis=5_000_000)
    m = IntervalMap.from_arrays(s, e)
    got = m.count_batch_np(qs, qe)
    ss, se = np.sort(s), np.sort(e)
    want = np.searchsorted(ss, qe, "right") - np.searchsorted(se, qs, "left")     # closed form, well-formed data
    assert np.array_equal(got.astype(np.int64), want.astype(np.int64))
    o = np.argsort(qs, kind="stable")                                              # position-sorted batch: no device sort
    got2 = m.count_batch_np(np.ascontiguousarray(qs[o]), np.ascontiguousarray(qe[o]))
    assert np.array_equal(got2, got[o])


@pytest.mark.parametrize("case,expect", [("random", 1), ("short_runs", 1), ("equal_keys", 1), ("long_run", 2), ("option_off", 2),
                                         ("presorted", 0)])
def test_narrow_sort_with_tie_fix_builds_the_same_arrays_as_the_composite_key(case, expect):
    """build() sorts by start alone (32-bit keys) and fixes equal starts in place; more than 16 equal starts fall back to
    the composite 64-bit key. Whatever path ran (siIndexLastSort), starts / ends / payload order (stable among fully
    equal keys, quirk Q3) / branch equal the oracle's."""
    import torch
    from oracle.pyoracle import Oracle
    from superintervals_b200.device import DeviceIndex, OPT_NARROW_SORT
    rng = np.random.default_rng(31)
    n = 60_000
    if case in ("random", "option_off", "presorted"):
        s = rng.integers(0, 5_000_000, n).astype(np.int32)
        e = (s + rng.integers(0, 4000, n)).astype(np.int32)
    elif case == "short_runs":        # every start 1..16 times, ends all over the place, negative coordinates too
        s = np.repeat(rng.choice(4_000_000, n // 8, replace=False) - 2_000_000, rng.integers(1, 17, n // 8)).astype(np.int32)   # distinct starts: no run above 16
        e = (s + rng.integers(0, 50_000, s.size)).astype(np.int32)
        p = rng.permutation(s.size); s, e = s[p], e[p]
    elif case == "equal_keys":        # runs whose ends repeat: insertion order must survive among equal (start, end)
        s = np.repeat(rng.choice(100_000, n // 6, replace=False), 6).astype(np.int32)
        e = (s + rng.integers(0, 3, s.size) * 100).astype(np.int32)
        p = rng.permutation(s.size); s, e = s[p], e[p]
    else:                              # one start 17 times among random data
        s = rng.integers(0, 5_000_000, n).astype(np.int32)
        s[rng.choice(n, 17, replace=False)] = 123_456
        e = (s + rng.integers(0, 4000, n)).astype(np.int32)
    if case == "presorted":
        o = np.lexsort((-e.astype(np.int64), s)); s, e = s[o], e[o]
    ix = DeviceIndex()
    if case == "option_off":
        ix.set_option(OPT_NARROW_SORT, 0)
    ix.build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    assert ix.last_sort() == expect
    orc = Oracle(s, e)
    gs, ge, gv, gb, gp = ix.export()
    assert np.array_equal(gs, orc.starts) and np.array_equal(ge, orc.ends)
    assert np.array_equal(gv, orc.data) and np.array_equal(gb, orc.branch)


@pytest.mark.parametrize("mode", ["cells_shuffled", "stream_sorted", "partitioned", "walk_fallback", "empty_index"])
def test_count_fanout_stores_every_count_into_the_extra_arrays(mode):
    """siCountFanoutDevice (the fused count + all-gather): besides its own output the count kernel stores every count at
    the same index of the extra arrays -- here two more arrays on the same GPU, at an offset like a slot of a gathered vector."""
    import torch
    from oracle.pyoracle import Oracle
    from superintervals_b200.device import (DeviceIndex, ORDER_SORTED, ORDER_UNSORTED, OPT_COUNT_ALGO, COUNT_WALK,
                                            OPT_CELLS_DIRECT_BYTES)
    s, e, qs, qe = W.config2(60_000, 70_001, 3, axis=3_000_000)
    if mode == "empty_index":
        s, e = s[:0], e[:0]
    order = ORDER_UNSORTED
    if mode == "stream_sorted":
        o = np.argsort(qs, kind="stable"); qs, qe = qs[o], qe[o]
        order = ORDER_SORTED
    ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    if mode == "walk_fallback":
        ix.set_option(OPT_COUNT_ALGO, COUNT_WALK)
    if mode == "partitioned":
        ix.set_option(OPT_CELLS_DIRECT_BYTES, 1)
    n = qs.size
    own = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    extra = torch.full((2, n + 64), -1, dtype=torch.int32, device="cuda")
    ptrs = [extra[0, 8:].data_ptr(), extra[1, 40:].data_ptr()]
    ix.count_fanout(torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda(), own, ptrs, order=order)
    torch.cuda.synchronize()
    want = Oracle(s, e).count_batch(qs, qe).astype(np.int64) if s.size else np.zeros(n, np.int64)
    assert np.array_equal(own.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(extra[0, 8:8 + n].cpu().numpy().astype(np.int64), want)
    assert np.array_equal(extra[1, 40:40 + n].cpu().numpy().astype(np.int64), want)
    assert (extra[0, :8] == -1).all() and (extra[0, 8 + n:] == -1).all() and (extra[1, :40] == -1).all() and (extra[1, 40 + n:] == -1).all()


def test_peer_barrier_kernel_signals_waits_and_times_out_instead_of_hanging():
    import ctypes as C
    import torch
    from superintervals_b200 import _lib
    L = _lib.lib()
    flags = torch.zeros(8, dtype=torch.int32, device="cuda")
    # a "peer" on the same GPU: the word signalled is the word awaited
    sig = (C.c_void_p * 2)(C.c_void_p(flags[0:].data_ptr()), C.c_void_p(flags[1:].data_ptr()))
    for seq in (1, 2, 3):
        assert L.siPeerBarrierDevice(sig, sig, 2, seq, flags[7:].data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert flags.cpu().tolist() == [3, 3, 0, 0, 0, 0, 0, 0]
