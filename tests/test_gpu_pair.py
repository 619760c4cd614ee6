"""GPU: pair cells (SI_OPT_PAIR_CELLS) -- both ranks of a coordinate cell in one 32-byte record, the table that
answers indexes whose rank cells do not fit L2. Counts must equal the oracle's (reference count(), hpp:651-825)
and the counts of the same index answered from the separate rank cells, for both record formats, with over-full
sides, malformed intervals, negative coordinates, clamped and inverted queries, and in the mixed-batch kernel."""
import zlib

import numpy as np
import pytest
import torch

from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _u32(t):
    return t.cpu().numpy().astype(np.uint32).astype(np.uint64)


def _build(s, e, mode):
    from superintervals_b200.device import DeviceIndex
    from superintervals_b200._lib import OPT_PAIR_CELLS
    ix = DeviceIndex()
    ix.set_option(OPT_PAIR_CELLS, mode)
    return ix.build(_dev(s), _dev(e))


def _intervals(rng, n, axis, lo_len, hi_len, origin=0):
    s = rng.integers(0, axis, n).astype(np.int64) + origin
    ln = rng.integers(lo_len, hi_len + 1, n)
    return s.astype(np.int32), (s + ln).astype(np.int32)


def _queries(rng, nq, axis, origin=0, stab_share=0.5, max_len=3000):
    """stabs, short ranges, long ranges, queries hanging over either end of the axis, a few inverted ones"""
    qs = rng.integers(-2000, axis + 2000, nq).astype(np.int64) + origin
    ln = np.where(rng.random(nq) < stab_share, 0, rng.integers(0, max_len, nq))
    qe = qs + ln
    inv = rng.random(nq) < 0.002
    qs, qe = np.where(inv, qe + 1, qs), np.where(inv, qs, qe)
    return qs.astype(np.int32), qe.astype(np.int32)


@pytest.mark.parametrize("name,n,axis,lens,want_fmt", [
    ("dense_c5", 200_000, 400_000, (150, 10_000), 4),       # 0.5 values per coordinate: four-bit offsets, cells of 16
    ("read_c2", 200_000, 5_000_000, (150, 10_000), 8),      # 0.04 per coordinate: one-byte offsets
    ("medium", 300_000, 1_000_000, (1, 50), 4),             # 0.3 per coordinate, short intervals
    ("thin", 100_000, 20_000_000, (10, 500), 8),            # 0.005 per coordinate: widest one-byte cell
])
def test_pair_cells_count_equals_oracle(name, n, axis, lens, want_fmt):
    from superintervals_b200.device import ORDER_UNSORTED
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    s, e = _intervals(rng, n, axis, *lens)
    qs, qe = _queries(rng, 300_000, axis)
    orc = Oracle(s, e)
    want = orc.count_batch(qs, qe)
    ix = _build(s, e, 2)
    info = ix.cells_info()["pair"]
    assert info["format"] == want_fmt, info
    got = _u32(ix.count(_dev(qs), _dev(qe), order=ORDER_UNSORTED))
    assert np.array_equal(got, want)
    plain = _build(s, e, 0)
    assert plain.cells_info()["pair"]["format"] == 0
    assert np.array_equal(_u32(plain.count(_dev(qs), _dev(qe), order=ORDER_UNSORTED)), want)


def test_pair_cells_default_is_off_for_an_index_that_fits_l2():
    rng = np.random.default_rng(7)
    s, e = _intervals(rng, 100_000, 2_000_000, 100, 2000)
    assert _build(s, e, 1).cells_info()["pair"]["format"] == 0


def test_pair_cells_overfull_sides_are_answered_from_the_arrays():
    """a few coordinates carrying 40 starts (and 40 ends) each: more than either format holds per side"""
    from superintervals_b200.device import ORDER_UNSORTED
    rng = np.random.default_rng(11)
    s, e = _intervals(rng, 150_000, 3_000_000, 100, 3000)
    hot = rng.integers(0, 3_000_000, 25)
    hs = np.repeat(hot, 40).astype(np.int32)
    he = (hs + 777).astype(np.int32)
    s, e = np.concatenate([s, hs]), np.concatenate([e, he])
    qs, qe = _queries(rng, 200_000, 3_000_000)
    # queries right on the hot coordinates and their ends
    qs = np.concatenate([qs, hs[:1000], he[:1000], hs[:1000] - 1, he[:1000] + 1]).astype(np.int32)
    qe = np.concatenate([qe, hs[:1000], he[:1000], hs[:1000] + 5, he[:1000] + 9]).astype(np.int32)
    ix = _build(s, e, 2)
    info = ix.cells_info()["pair"]
    assert info["format"] == 8 and info["overfull"] >= 50, info
    assert np.array_equal(_u32(ix.count(_dev(qs), _dev(qe), order=ORDER_UNSORTED)), Oracle(s, e).count_batch(qs, qe))


def test_pair_cells_with_malformed_intervals_and_negative_coordinates():
    from superintervals_b200.device import ORDER_UNSORTED
    rng = np.random.default_rng(13)
    origin = -1_500_000
    s, e = _intervals(rng, 200_000, 3_000_000, 50, 5000, origin=origin)
    s[1000], e[1000] = 5000, 4000            # start > end (quirk Q6): tested per query from the side list
    s[77], e[77] = -20_000, -30_000
    qs, qe = _queries(rng, 200_000, 3_000_000, origin=origin)
    ix = _build(s, e, 2)
    assert ix.cells_info()["pair"]["format"] == 8
    assert np.array_equal(_u32(ix.count(_dev(qs), _dev(qe), order=ORDER_UNSORTED)), Oracle(s, e).count_batch(qs, qe))


def test_pair_cells_extreme_coordinates():
    from superintervals_b200.device import ORDER_UNSORTED
    rng = np.random.default_rng(17)
    s = rng.integers(2**31 - 400_000, 2**31 - 5000, 150_000).astype(np.int64)
    e = np.minimum(s + rng.integers(0, 4000, s.size), 2**31 - 1)
    s, e = s.astype(np.int32), e.astype(np.int32)
    qs = rng.integers(2**31 - 420_000, 2**31 - 1, 100_000).astype(np.int64)
    qe = np.minimum(qs + rng.integers(0, 300, qs.size), 2**31 - 1)
    qs, qe = qs.astype(np.int32), qe.astype(np.int32)
    qs[:3], qe[:3] = [-2**31, 2**31 - 1, -2**31], [2**31 - 1, 2**31 - 1, -2**31]
    ix = _build(s, e, 2)
    assert ix.cells_info()["pair"]["format"] != 0
    assert np.array_equal(_u32(ix.count(_dev(qs), _dev(qe), order=ORDER_UNSORTED)), Oracle(s, e).count_batch(qs, qe))


def test_pair_cells_search_values_and_sorted_batches_unchanged():
    """the fill and the streaming kernel keep reading the separate tables: same lists, same counts"""
    from superintervals_b200.device import ORDER_SORTED, ORDER_UNSORTED
    rng = np.random.default_rng(19)
    s, e = _intervals(rng, 200_000, 400_000, 150, 3000)
    qs, qe = _queries(rng, 100_000, 400_000, stab_share=0.3, max_len=500)
    keep = qs <= qe
    qs, qe = qs[keep], qe[keep]
    orc = Oracle(s, e)
    ix = _build(s, e, 2)
    off_o, res = orc.search_batch(qs, qe, want=("values",))
    off, vals = ix.search_values(_dev(qs), _dev(qe), order=ORDER_UNSORTED)
    assert np.array_equal(off.cpu().numpy().astype(np.uint64), off_o) and np.array_equal(vals.cpu().numpy(), res["values"])
    o = np.argsort(qs, kind="stable")
    assert np.array_equal(_u32(ix.count(_dev(qs[o]), _dev(qe[o]), order=ORDER_SORTED)), orc.count_batch(qs[o], qe[o]))


def test_mixed_batch_over_contigs_with_pair_cells():
    """mode B in one launch (examples/bed-intersect-si.rs:100-123: one map per contig) with every contig on pair cells"""
    from superintervals_b200.genome import GenomeIndex
    rng = np.random.default_rng(23)
    axes = [400_000, 2_000_000, 150_000, 900_000]
    ns = [120_000, 90_000, 60_000, 30_000]
    g = GenomeIndex([f"c{i}" for i in range(4)], ns, rank=0, world=1, pair_cells=2)
    plain = GenomeIndex([f"c{i}" for i in range(4)], ns, rank=0, world=1, pair_cells=0)
    orcs = []
    for c, (n, axis) in enumerate(zip(ns, axes)):
        s, e = _intervals(rng, n, axis, 100, 4000)
        g.build_contig(c, _dev(s), _dev(e))
        plain.build_contig(c, _dev(s), _dev(e))
        assert g.index(c).cells_info()["pair"]["format"] != 0
        orcs.append(Oracle(s, e))
    nq = 300_000
    cid = rng.integers(0, 4, nq).astype(np.int32)
    qs, qe = _queries(rng, nq, 2_000_000)
    want = np.zeros(nq, np.uint64)
    for c in range(4):
        m = cid == c
        want[m] = orcs[c].count_batch(qs[m], qe[m])
    got = g.count_mixed(_dev(cid), _dev(qs), _dev(qe))
    assert np.array_equal(_u32(got), want)
    assert np.array_equal(_u32(plain.count_mixed(_dev(cid), _dev(qs), _dev(qe))), want)
