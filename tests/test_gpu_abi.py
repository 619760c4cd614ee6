"""GPU: the C ABI as a drop-in -- golden fixtures from the real reference, the reference's
unit-test vectors through the single-query entry points, and the same ctypes driver run
against the reference's own C library (oracle/_ref/libsi_cref.so) and ours."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import canonical_values, golden_files, load_vectors
from superintervals_b200 import IntervalMap, _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("path", golden_files("presorted"), ids=lambda p: p.split("/")[-1])
def test_golden_presorted_bit_exact(path):
    g = np.load(path)
    m = IntervalMap.from_arrays(g["in_starts"], g["in_ends"])
    assert np.array_equal(m.starts, g["starts"]) and np.array_equal(m.ends, g["ends"])
    assert np.array_equal(m.data_index, g["data"]) and np.array_equal(m.branch, g["branch"])
    qs, qe = g["qs"], g["qe"]
    assert np.array_equal(m.count_batch_np(qs, qe), g["count"])
    assert np.array_equal(m.has_overlaps_batch(qs, qe), g["has_overlaps"])
    off, vals = m.search_values_batch_csr(qs, qe)
    assert np.array_equal(off, g["offsets"]) and np.array_equal(vals, g["values"])
    assert np.array_equal(vals, g["values_large"])           # search_values_large: same output (hpp:588)
    off, keys = m.search_keys_batch_csr(qs, qe)
    assert np.array_equal(off, g["offsets"]) and np.array_equal(keys, g["keys"])
    off, idx = m.search_idxs_batch_csr(qs, qe)
    for q in range(len(qs)):                                  # same set as the C++ overload (Q2 reorders it)
        a, b = int(off[q]), int(off[q + 1])
        assert np.array_equal(np.sort(idx[a:b]), np.sort(g["idxs_cpp"][a:b].astype(np.int64)))
    cnt, cov = m.coverage_batch(qs, qe)
    assert np.array_equal(cnt, g["cov_count"]) and np.array_equal(cov, g["cov_sum"])


@pytest.mark.parametrize("path", golden_files("shuffled"), ids=lambda p: p.split("/")[-1])
def test_golden_shuffled(path):
    g = np.load(path)
    m = IntervalMap.from_arrays(g["in_starts"], g["in_ends"])
    assert np.array_equal(m.starts, g["starts"]) and np.array_equal(m.ends, g["ends"])
    assert np.array_equal(m.branch, g["branch"])
    assert np.array_equal(m.count_batch_np(g["qs"], g["qe"]), g["count"])
    off, vals = m.search_values_batch_csr(g["qs"], g["qe"])
    _, keys = m.search_keys_batch_csr(g["qs"], g["qe"])
    assert np.array_equal(off, g["offsets"]) and np.array_equal(keys, g["keys"])
    assert np.array_equal(canonical_values(off, vals, keys), canonical_values(g["offsets"], g["values"], g["keys"]))


def _drive(L, case):
    """Run one tests.cpp-style case through a library exporting the reference C ABI. Returns a
    transcript that must be identical for the reference's library and ours."""
    si = L.createSuperIntervals()
    out = []
    for s, e, d in case["intervals"]:
        L.addInterval(si, s, e, d)
    L.indexSuperIntervals(si)
    c = si.contents
    out.append(("size", int(L.sizeSuperIntervals(si))))
    out.append(("sorted", [(c.starts[i], c.ends[i], c.data[i]) for i in range(c.size)]))
    if c.size:
        out.append(("branch", [int(c.branch[i]) for i in range(c.size)]))
    qlist = [tuple(q["q"]) for q in case.get("queries", [])]
    if "batch" in case:
        qlist += list(zip(case["batch"]["starts"], case["batch"]["ends"]))
    for s, e in qlist:
        out.append(("any", bool(L.anyOverlaps(si, s, e))))
        out.append(("count", int(L.countOverlaps(si, s, e))))
        out.append(("ub", int(L.upperBound(si, e)), int(c.idx)))
        r = L.createIndexResult()
        L.searchValues(si, s, e, C.byref(r))
        out.append(("values", [r.data[i] for i in range(r.size)]))
        L.searchValues(si, s, e, C.byref(r))                  # appends (ref c.h:209)
        out.append(("values_appended", int(r.size)))
        L.clearIndexResult(C.byref(r))
        L.searchIdxs(si, s, e, C.byref(r))
        out.append(("idxs", [r.data[i] for i in range(r.size)]))
        L.clearIndexResult(C.byref(r))
        L.searchPoint(si, s, C.byref(r))
        out.append(("point", [r.data[i] for i in range(r.size)]))
        L.destroyIndexResult(C.byref(r))
        k = L.createKeyResult()
        L.searchKeys(si, s, e, C.byref(k))
        out.append(("keys", [(k.data[i].start, k.data[i].end) for i in range(k.size)]))
        L.destroyKeyResult(C.byref(k))
        it = L.createItemResult()
        L.searchItems(si, s, e, C.byref(it))
        out.append(("items", [(it.data[i].start, it.data[i].end, it.data[i].data) for i in range(it.size)]))
        L.destroyItemResult(C.byref(it))
        cnt, cov = C.c_size_t(0), C.c_int32(0)
        L.coverage(si, s, e, C.byref(cnt), C.byref(cov))
        out.append(("coverage", int(cnt.value), int(cov.value)))
        buf = (C.c_int32 * 64)()
        nfound = C.c_size_t(0)
        L.findOverlaps(si, s, e, buf, C.byref(nfound))
        out.append(("find", [buf[i] for i in range(nfound.value)]))
    L.destroySuperIntervals(si)
    return out


@pytest.mark.parametrize("case", load_vectors(), ids=lambda c: c["name"].split(" (")[0])
def test_reference_unit_vectors_through_c_abi(case):
    L = _lib.lib()
    L.si_b200_clear_error()
    t = dict((x[0] + str(i), x[1:]) for i, x in enumerate(_drive(L, case)))
    _lib.check(case["name"])
    got = _drive(L, case)
    it = iter(got[3:] if case["intervals"] else got[2:])
    for q in case.get("queries", []):
        rec = {}
        for _ in range(11):
            x = next(it)
            rec[x[0]] = x[1:] if len(x) > 2 else x[1]
        op = q["op"]
        if op == "count": assert rec["count"] == q["expect"]
        if op == "has_overlaps": assert rec["any"] == q["expect"]
        if op == "search_values":
            if "expect" in q: assert rec["values"] == q["expect"]
            if "expect_size" in q: assert len(rec["values"]) == q["expect_size"] and rec["values_appended"] == 2 * q["expect_size"]
            if "expect_last" in q: assert rec["values"][-1] == q["expect_last"]
            assert rec["find"] == rec["values"]
        if op == "search_idxs": assert rec["idxs"] == q["expect"]
        if op == "search_keys": assert [list(k) for k in rec["keys"]] == q["expect"]
        if op == "search_items": assert [list(k) for k in rec["items"]] == q["expect"]
        if op == "coverage":
            if "expect_count" in q: assert rec["coverage"][0] == q["expect_count"]
            assert rec["coverage"][1] == q["expect_sum"]
    if "expect_branch" in case:
        want = [(1 << 64) - 1 if v < 0 else v for v in case["expect_branch"]]
        assert got[2][1] == want
    del t


CREF = os.path.join(ROOT, "oracle", "_ref", "libsi_cref.so")


@pytest.mark.skipif(not os.path.exists(CREF), reason="reference C library not built")
@pytest.mark.parametrize("case", [c for c in load_vectors() if c["intervals"]], ids=lambda c: c["name"].split(" (")[0])
def test_same_driver_reference_c_library_vs_ours(case):
    """The identical ctypes transcript from the reference's compiled c_superintervals.h and from
    libsuperintervals_b200.so (tie order: these cases are inserted pre-sorted or tie-free except
    'duplicates', whose equal keys keep insertion order under glibc qsort's merge sort too)."""
    ref = _lib.bind(C.CDLL(CREF))
    ours = _lib.lib()
    assert _drive(ref, case) == _drive(ours, case)


def test_python_doc_example():
    """src/superintervals/README.md:16-26 -- values ride as payload objects on the host."""
    m = IntervalMap.from_arrays([10, 15, 30], [20, 25, 40], ["A", "B", "C"])
    assert m.count_batch(np.array([5, 18, 35], np.int32), np.array([12, 22, 45], np.int32)) == [1, 2, 1]
    assert m.search_values_batch(np.array([5, 18, 35], np.int32), np.array([12, 22, 45], np.int32)) == [["A"], ["B", "A"], ["C"]]
    # the reference's Python class routes idxs / keys / items through the C++ vector overload: first run ASCENDING (Q2)
    assert m.search_idxs_batch(np.array([5, 18, 35], np.int32), np.array([12, 22, 45], np.int32)) == [[0], [0, 1], [2]]
    assert m.search_values(8, 20) == ["B", "A"] and m.count(8, 20) == 2 and m.has_overlaps(8, 20)
    assert m.search_keys(8, 20) == [(10, 20), (15, 25)] and m.search_items(31, 31) == [(30, 40, "C")]
    assert m.coverage(12, 18) == (2, 9) and m.at(1) == (15, 25, "B")


@pytest.mark.parametrize("name", ["nested", "reads", "readme"])
def test_python_query_lists_come_in_the_reference_modules_order(name):
    """tests/golden/py_queries.json: the UNMODIFIED reference Cython module's answers (tools/make_golden_py_queries.py).
    search_idxs / search_keys / search_items / search_idxs_batch go through the C++ vector overload (pyx:299,314,335,440 ->
    hpp:879-905: first contiguous run ascending, branch-walk hits descending); search_values(_batch) are all-descending."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "py_queries.json")))[name]
    m = IntervalMap.from_arrays(g["starts"], g["ends"], list(range(len(g["starts"]))))
    qs, qe = np.array(g["qs"], np.int32), np.array(g["qe"], np.int32)
    assert m.search_idxs_batch(qs, qe) == g["search_idxs_batch"]
    assert m.search_values_batch(qs, qe) == g["search_values_batch"]
    assert m.count_batch(qs, qe) == g["count_batch"]
    assert [bool(x) for x in m.has_overlaps_batch(qs, qe)] == g["has_overlaps"]
    for k in range(0, len(qs), 3):
        a, b = int(qs[k]), int(qe[k])
        assert m.search_idxs(a, b) == g["search_idxs"][k]
        assert [list(x) for x in m.search_keys(a, b)] == g["search_keys"][k]
        assert [list(x) for x in m.search_items(a, b)] == g["search_items"][k]
        assert m.search_values(a, b) == g["search_values"][k]
        assert list(m.coverage(a, b)) == g["coverage"][k] and m.has_overlaps(a, b) == g["has_overlaps"][k]


def test_rebuild_after_adding_and_clear():
    m = IntervalMap()
    for i in range(50):
        m.add(100 - i, 200 - i, i)                         # reverse order: forces the device sort
    m.build()
    assert m.count(0, 1000) == 50 and m.starts.tolist() == sorted(m.starts.tolist())
    m.add(500, 600, "late")
    m.build()
    assert m.count(550, 560) == 1 and m.search_values(550, 560) == ["late"] and m.count(0, 1000) == 51
    m.clear()
    m.build()
    assert m.count(0, 1000) == 0 and not m.has_overlaps(0, 1000)


@pytest.mark.parametrize("resident", [False, True])
def test_single_query_calls_use_the_mailbox_and_long_lists_fall_back(resident):
    """The one-query-per-call C functions (c_superintervals.h:537-821) answer through the mapped pinned mailbox (one
    launch per call, or -- SI_OPT_RESIDENT_QUERIES -- a resident polling warp; no copies); a result list longer than the
    mailbox takes the batch path. Same answers as the oracle either way."""
    import ctypes as C
    from superintervals_b200 import _lib
    from oracle.pyoracle import Oracle
    L = _lib.lib()
    rng = np.random.default_rng(9)
    n = 120_000
    s = rng.integers(0, 1000, n).astype(np.int32)                 # everything overlaps everything: lists of ~n hits
    e = (s + rng.integers(5000, 9000, n)).astype(np.int32)
    orc = Oracle(s, e)
    si = L.createSuperIntervals()
    L.addIntervals(si, s.ctypes.data, e.ctypes.data, None, n)
    L.indexSuperIntervals(si)
    _lib.check("indexSuperIntervals")
    assert L.siIndexSetOption(L.siIndexOf(si), _lib.OPT_RESIDENT_QUERIES, 1 if resident else 0) == 0
    for qs, qe in ((2000, 3000), (0, 10), (999, 999), (9500, 9999), (20000, 30000), (500, 400)):
        a, b = np.array([qs], np.int32), np.array([qe], np.int32)
        want = int(orc.count_batch(a, b)[0])
        assert int(L.countOverlaps(si, qs, qe)) == want
        _, res = orc.search_batch(a, b, want=("values", "keys", "idxs"))
        r = L.createIndexResult()
        L.searchValues(si, qs, qe, C.byref(r))
        got = np.ctypeslib.as_array(r.data, shape=(int(r.size),)).copy() if r.size else np.zeros(0, np.int32)
        assert np.array_equal(got, res["values"]), (qs, qe)
        L.searchValues(si, qs, qe, C.byref(r))                    # appends (quirk Q4)
        assert int(r.size) == 2 * want
        L.destroyIndexResult(C.byref(r))
        k = L.createKeyResult()
        L.searchKeys(si, qs, qe, C.byref(k))
        assert int(k.size) == want
        if want:
            kk = np.ctypeslib.as_array(C.cast(k.data, C.POINTER(C.c_int32)), shape=(want, 2))
            assert np.array_equal(kk, res["keys"])
        L.destroyKeyResult(C.byref(k))
        cnt, cov = C.c_size_t(0), C.c_int32(0)
        L.coverage(si, qs, qe, C.byref(cnt), C.byref(cov))
        assert int(cnt.value) == want
        assert bool(L.anyOverlaps(si, qs, qe)) == bool(orc.has_overlaps_batch(a, b)[0])
    assert int(L.upperBound(si, 500)) == int(np.searchsorted(orc.starts, 500, "right")) - 1
    assert int(L.upperBound(si, -5)) == 2**64 - 1
    _lib.check("single queries")
    L.destroySuperIntervals(si)


def test_resident_query_kernel_survives_idle_gaps_rebuilds_and_a_device_synchronise():
    """SI_OPT_RESIDENT_QUERIES: the polling warp leaves after 0.2 ms without work (2 ms at most) and is relaunched on demand;
    a rebuild stops it first (it reads the old arrays); a device-wide synchronise returns promptly while it is resident."""
    import ctypes as C
    import time
    import torch
    from superintervals_b200 import _lib, workloads as W
    from oracle.pyoracle import Oracle
    L = _lib.lib()
    s, e, qs, qe = W.config3(30_000, 400, 5, axis=2_000_000)
    orc = Oracle(s, e)
    want = orc.count_batch(qs, qe)
    si = L.createSuperIntervals()
    L.addIntervals(si, s.ctypes.data, e.ctypes.data, None, s.size)
    L.indexSuperIntervals(si)
    L.siIndexSetOption(L.siIndexOf(si), _lib.OPT_RESIDENT_QUERIES, 1)
    for k in range(qs.size):
        assert int(L.countOverlaps(si, int(qs[k]), int(qe[k]))) == int(want[k])
        if k % 97 == 0:
            time.sleep(0.004)                                     # the kernel has left by now: the next call relaunches it
        if k == 150:
            t0 = time.perf_counter(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            assert dt < 0.5                                       # bounded by the kernel's 2 ms lifetime, not by the caller's loop
    _, res = orc.search_batch(qs[:50], qe[:50], want=("values",))
    r = L.createIndexResult()
    for k in range(50):
        L.searchValues(si, int(qs[k]), int(qe[k]), C.byref(r))
    got = np.ctypeslib.as_array(r.data, shape=(int(r.size),)).copy() if r.size else np.zeros(0, np.int32)
    assert np.array_equal(got, res["values"])
    L.destroyIndexResult(C.byref(r))
    # add + rebuild while the resident kernel may still be polling
    L.addInterval(si, 10, 1_999_999, 777)
    L.indexSuperIntervals(si)
    s2, e2 = np.append(s, np.int32(10)), np.append(e, np.int32(1_999_999))
    want2 = Oracle(s2, e2).count_batch(qs[:100], qe[:100])
    for k in range(100):
        assert int(L.countOverlaps(si, int(qs[k]), int(qe[k]))) == int(want2[k])
    _lib.check("resident single queries")
    L.destroySuperIntervals(si)                                   # stops the kernel, frees the mailbox


def test_count_batch_32_bit_counts_equal_the_size_t_call():
    from superintervals_b200 import _lib, workloads as W
    L = _lib.lib()
    s, e = W.config2_intervals(100_000, 5, axis=3_000_000)
    qs, qe = W.config2_queries(14_000_000, 5, axis=3_000_000)      # above the 12 M pipeline threshold
    si = L.createSuperIntervals()
    L.addIntervals(si, s.ctypes.data, e.ctypes.data, None, s.size)
    L.indexSuperIntervals(si)
    for m in (1, 1000, qs.size):
        c64, c32 = np.zeros(m, np.uint64), np.zeros(m, np.uint32)
        L.countOverlapsBatch(si, qs.ctypes.data, qe.ctypes.data, m, c64.ctypes.data)
        L.countOverlapsBatch32(si, qs.ctypes.data, qe.ctypes.data, m, c32.ctypes.data)
        _lib.check("countOverlapsBatch32")
        assert np.array_equal(c64, c32.astype(np.uint64))
    ss, se = np.sort(s), np.sort(e)
    assert np.array_equal(c32.astype(np.int64), np.searchsorted(ss, qe, "right") - np.searchsorted(se, qs, "left"))
    L.destroySuperIntervals(si)
