"""CPU: pin the oracle (oracle/si_oracle.c) against
  1. golden fixtures produced by the REAL reference (tests/golden/*.npz, tools/make_golden.py),
  2. the reference's own unit-test vectors (tests/golden/tests_cpp_vectors.json),
  3. a brute-force O(N*Q) checker and the closed-form rank count,
  4. the real reference itself where oracle/_ref is present (this container and the GPU box).
"""
import numpy as np
import pytest

from helpers import NONE64, brute_force, canonical_values, csr_from_hits, golden_files, load_vectors
from oracle.pyoracle import Oracle, Reference
from superintervals_b200 import workloads as W


@pytest.mark.parametrize("path", golden_files("presorted"), ids=lambda p: p.split("/")[-1])
def test_oracle_matches_reference_golden_presorted(path):
    g = np.load(path)
    o = Oracle(g["in_starts"], g["in_ends"])
    assert np.array_equal(o.starts, g["starts"]) and np.array_equal(o.ends, g["ends"])
    assert np.array_equal(o.data, g["data"])
    assert np.array_equal(o.branch, g["branch"])
    qs, qe = g["qs"], g["qe"]
    c = o.count_batch(qs, qe)
    assert np.array_equal(c, g["count"])
    assert np.array_equal(c, g["count_linear"])          # hpp:623: same answer, different walk
    assert np.array_equal(o.has_overlaps_batch(qs, qe), g["has_overlaps"])
    off, res = o.search_batch(qs, qe, want=("values", "idxs", "keys"))
    assert np.array_equal(off, g["offsets"])
    assert np.array_equal(res["values"], g["values"])    # exact order: descending position
    assert np.array_equal(res["keys"], g["keys"])
    # C++ vector search_idxs puts its first run ascending (Q2): same set, per query
    for q in range(len(qs)):
        a, b = int(off[q]), int(off[q + 1])
        assert np.array_equal(np.sort(res["idxs"][a:b]), np.sort(g["idxs_cpp"][a:b]))


@pytest.mark.parametrize("path", golden_files("shuffled"), ids=lambda p: p.split("/")[-1])
def test_oracle_matches_reference_golden_shuffled(path):
    g = np.load(path)
    o = Oracle(g["in_starts"], g["in_ends"])
    assert np.array_equal(o.starts, g["starts"]) and np.array_equal(o.ends, g["ends"])
    assert np.array_equal(o.branch, g["branch"])
    assert np.array_equal(o.count_batch(g["qs"], g["qe"]), g["count"])
    off, res = o.search_batch(g["qs"], g["qe"], want=("values", "keys"))
    assert np.array_equal(off, g["offsets"]) and np.array_equal(res["keys"], g["keys"])
    # payload order inside exact-duplicate (start,end) groups is std::sort's (unstable): canonicalise
    assert np.array_equal(canonical_values(off, res["values"], res["keys"]),
                          canonical_values(g["offsets"], g["values"], g["keys"]))


def test_count_large_differs_only_on_malformed():
    """count_large (hpp:834) assumes start <= end; on well-formed fixtures it equals count."""
    for path in golden_files("presorted"):
        g = np.load(path)
        if (g["in_starts"] <= g["in_ends"]).all() and (g["qs"] <= g["qe"]).all():
            assert np.array_equal(g["count"], g["count_large"]), path


def _run_case(case):
    iv = np.array(case["intervals"], np.int32).reshape(-1, 3)
    return Oracle(iv[:, 0], iv[:, 1], iv[:, 2]), iv


@pytest.mark.parametrize("case", load_vectors(), ids=lambda c: c["name"].split(" (")[0])
def test_oracle_reference_unit_vectors(case):
    o, iv = _run_case(case)
    for q in case.get("queries", []):
        s, e = q["q"]
        op = q["op"]
        if op == "count":
            assert int(o.count_batch([s], [e])[0]) == q["expect"]
        elif op == "has_overlaps":
            assert bool(o.has_overlaps_batch([s], [e])[0]) == q["expect"]
        elif op in ("search_values", "search_idxs", "search_keys", "search_items"):
            _, res = o.search_batch([s], [e], want=("values", "idxs", "keys"))
            if op == "search_values":
                if "expect" in q: assert res["values"].tolist() == q["expect"]
                if "expect_size" in q: assert len(res["values"]) == q["expect_size"]
                if "expect_last" in q: assert res["values"][-1] == q["expect_last"]
            elif op == "search_idxs":
                assert res["idxs"].tolist() == q["expect"]
            elif op == "search_keys":
                assert res["keys"].tolist() == q["expect"]
            else:
                got = [[int(a), int(b), int(v)] for (a, b), v in zip(res["keys"], res["values"])]
                assert got == q["expect"]
        elif op == "coverage":
            hit = brute_force(o.starts, o.ends, [s], [e])[0]
            cov = int((np.minimum(o.ends[hit], e).astype(np.int64) - np.maximum(o.starts[hit], s)).sum())
            if "expect_count" in q: assert int(hit.sum()) == q["expect_count"]
            assert cov == q["expect_sum"]
    if "batch" in case:
        b = case["batch"]
        assert o.count_batch(b["starts"], b["ends"]).tolist() == b["count_batch"]
        off, res = o.search_batch(b["starts"], b["ends"])
        got = [res["values"][int(off[i]):int(off[i + 1])].tolist() for i in range(len(b["starts"]))]
        assert got == b["search_values_batch"]
    if "expect_branch" in case:
        want = np.array([NONE64 if v < 0 else v for v in case["expect_branch"]], np.uint64)
        assert np.array_equal(o.branch, want)


@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_brute_force(seed):
    rng = np.random.default_rng(seed)
    n, nq = int(rng.integers(1, 400)), 300
    s = rng.integers(-50, 200, n).astype(np.int32)
    e = (s + rng.integers(-3, 60, n)).astype(np.int32)          # includes malformed start > end
    qs = rng.integers(-60, 260, nq).astype(np.int32)
    qe = (qs + rng.integers(-5, 80, nq)).astype(np.int32)
    o = Oracle(s, e)
    # sortedness contract of build()
    assert (np.diff(o.starts.astype(np.int64)) >= 0).all()
    same = np.diff(o.starts.astype(np.int64)) == 0
    assert (np.diff(o.ends.astype(np.int64))[same] <= 0).all()
    hit = brute_force(o.starts, o.ends, qs, qe)
    off, idx = csr_from_hits(hit)
    assert np.array_equal(o.count_batch(qs, qe), hit.sum(1).astype(np.uint64))
    off_o, res = o.search_batch(qs, qe, want=("idxs",))
    assert np.array_equal(off_o, off) and np.array_equal(res["idxs"], idx)
    # branch definition (hpp:117-129): nearest previous end >= mine
    ends = o.ends.astype(np.int64)
    for i in range(n):
        prev = np.flatnonzero(ends[:i] >= ends[i])
        want = NONE64 if prev.size == 0 else np.uint64(prev[-1])
        assert o.branch[i] == want


def test_closed_form_rank_count():
    """count == #{starts <= qe} - #{ends < qs} on well-formed data."""
    s, e, qs, qe = W.config1(50_000, 0)
    o = Oracle(s, e)
    want = np.searchsorted(np.sort(s), qe, "right") - np.searchsorted(np.sort(e), qs, "left")
    assert np.array_equal(o.count_batch(qs, qe), want.astype(np.uint64))


def test_upper_bound_semantics():
    o = Oracle([5, 5, 9, 20], [6, 5, 30, 21])
    assert o.upper_bound(4) == (1 << 64) - 1      # SIZE_MAX: nothing starts at or before 4
    assert o.upper_bound(5) == 1 and o.upper_bound(8) == 1 and o.upper_bound(9) == 2 and o.upper_bound(10**9) == 3


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("gen", ["c1", "c2", "c3"])
def test_oracle_equals_live_reference(gen):
    s, e, qs, qe = {"c1": lambda: W.config1(100_000, 7), "c2": lambda: W.config2(20_000, 60_000, 8, axis=500_000),
                    "c3": lambda: W.config3(60_000, 60_000, 9, axis=4_000_000)}[gen]()
    ss, se = W.sort_by_start(s, e)               # start-sorted input as `bedtools sort` leaves it
    for a, b in ((s, e), (ss, se)):
        o, r = Oracle(a, b), Reference(a, b)
        rs, re_, rd, rb = r.export()
        assert np.array_equal(o.starts, rs) and np.array_equal(o.ends, re_) and np.array_equal(o.branch, rb)
        assert np.array_equal(o.count_batch(qs, qe), r.count_batch(qs, qe, 0, threads=2))
        off, res = o.search_batch(qs[:5000], qe[:5000], want=("values", "keys"))
        roff, rvals = r.search_values_batch(qs[:5000], qe[:5000])
        _, rkeys = r.search_keys_batch(qs[:5000], qe[:5000])
        assert np.array_equal(off, roff) and np.array_equal(res["keys"], rkeys)
        assert np.array_equal(canonical_values(off, res["values"], res["keys"]), canonical_values(roff, rvals, rkeys))
