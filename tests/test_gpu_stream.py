"""GPU: the streaming count kernel (csrc/stream_kernels.cuh) -- position-sorted batches answered from
TMA-staged rank bits -- against the oracle, and against the rank-cells kernel on every input shape that
takes one of its side doors (unsorted tiles, inverted queries, triple coordinates, ragged tails,
unaligned pointers, sparse indexes without rank bits)."""
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


def _kernels(ix):
    return {k for k, _ in ix.read_timings()}


@pytest.fixture(scope="module")
def dense():
    """C2-shaped: read-length intervals, one per ~25 coordinates."""
    from superintervals_b200.device import DeviceIndex
    rng = np.random.default_rng(7)
    n, axis = 200_000, 5_000_000
    ln = np.exp(rng.uniform(np.log(150), np.log(10_000), n)).astype(np.int64)
    s = rng.integers(0, axis - ln).astype(np.int32)
    e = (s + ln - 1).astype(np.int32)
    nq = 600_001                                    # ragged: not a multiple of the 2048-query tile; ~0.12 queries per coordinate
    lq = np.exp(rng.uniform(0, np.log(10_000), nq)).astype(np.int64)
    qs = rng.integers(0, axis - lq).astype(np.int32)
    qe = (qs + lq - 1).astype(np.int32)
    ix = DeviceIndex().build(_dev(s), _dev(e))
    return s, e, qs, qe, ix, Oracle(s, e)


def test_rank_bits_are_built_on_a_dense_index(dense):
    info = dense[4].bits_info()
    assert info["built"] and info["words"] > 0 and info["bytes"] == 12 * info["words"]


def test_sorted_batch_streams_and_matches_the_oracle(dense):
    from superintervals_b200.device import OPT_TIMING, ORDER_AUTO, ORDER_SORTED
    s, e, qs, qe, ix, orc = dense
    o = np.argsort(qs, kind="stable")
    sq, se = qs[o], qe[o]
    want = orc.count_batch(sq, se)
    ix.set_option(OPT_TIMING, 1)
    got = _u32(ix.count(_dev(sq), _dev(se), order=ORDER_SORTED))
    assert _kernels(ix) == {"count_stream"}
    assert np.array_equal(got, want)
    tiles, back = ix.stream_stats()
    assert tiles == (sq.size + 2047) // 2048 and back == 0            # every tile answered from its TMA-staged window
    got = _u32(ix.count(_dev(sq), _dev(se), order=ORDER_AUTO))       # found sorted by the device check
    assert _kernels(ix) == {"count_stream"}
    ix.set_option(OPT_TIMING, 0)
    assert np.array_equal(got, want)


def test_shuffled_batch_keeps_the_cells_kernel_unless_forced(dense):
    from superintervals_b200.device import OPT_STREAM, OPT_TIMING, ORDER_AUTO, ORDER_SORTED, ORDER_UNSORTED
    s, e, qs, qe, ix, orc = dense
    want = orc.count_batch(qs, qe)
    dqs, dqe = _dev(qs), _dev(qe)
    ix.set_option(OPT_TIMING, 1)
    for order in (ORDER_AUTO, ORDER_UNSORTED):
        assert np.array_equal(_u32(ix.count(dqs, dqe, order=order)), want)
        assert _kernels(ix) == {"count_cells"}
    # a caller who wrongly promises a sorted batch still gets exact counts: tiles whose window does not fit
    # the staging buffers go to the rank-cells code
    assert np.array_equal(_u32(ix.count(dqs, dqe, order=ORDER_SORTED)), want)
    assert _kernels(ix) == {"count_stream"}
    tiles, back = ix.stream_stats()
    assert back == tiles                                              # shuffled: no tile's window fits
    ix.set_option(OPT_STREAM, 2)
    assert np.array_equal(_u32(ix.count(dqs, dqe, order=ORDER_UNSORTED)), want)
    assert _kernels(ix) == {"count_stream"}
    ix.set_option(OPT_STREAM, 0)
    o = np.argsort(qs, kind="stable")
    assert np.array_equal(_u32(ix.count(_dev(qs[o]), _dev(qe[o]), order=ORDER_SORTED)), want[o])
    assert _kernels(ix) == {"count_cells"}
    ix.set_option(OPT_STREAM, 1)
    ix.set_option(OPT_TIMING, 0)


def test_inverted_and_out_of_span_queries_in_a_sorted_batch(dense):
    from superintervals_b200.device import ORDER_SORTED
    s, e, qs, qe, ix, orc = dense
    o = np.argsort(qs, kind="stable")
    sq, se = qs[o].copy(), qe[o].copy()
    rng = np.random.default_rng(3)
    k = rng.choice(sq.size, 40, replace=False)
    se[k] = sq[k] - rng.integers(1, 3000, k.size).astype(np.int32)     # qs > qe (quirk Q6): the walk's definition
    sq[:300] = np.int32(-2_000_000_000); se[:300] = np.int32(-1_999_999_000)   # far below the index
    sq[-300:] = np.int32(2_000_000_000); se[-300:] = np.int32(2_100_000_000)   # far above
    sq[300] = np.iinfo(np.int32).min; se[300] = np.iinfo(np.int32).max         # everything
    assert np.array_equal(_u32(ix.count(_dev(sq), _dev(se), order=ORDER_SORTED)), orc.count_batch(sq, se))
    tiles, back = ix.stream_stats()
    assert 0 < back < tiles                                           # tiles holding an inverted query went to the walk


def test_unaligned_pointers_and_tiny_batches(dense):
    from superintervals_b200.device import OPT_STREAM, ORDER_SORTED
    s, e, qs, qe, ix, orc = dense
    o = np.argsort(qs, kind="stable")
    sq, se = qs[o], qe[o]
    dqs, dqe = _dev(sq), _dev(se)
    out = torch.empty(sq.size, dtype=torch.int32, device="cuda")
    for a, b in ((1, 50_001), (3, 4099), (7, 8), (5, 2053)):
        ix.count(dqs[a:b], dqe[a:b], out=out[a:b], order=ORDER_SORTED)
        assert np.array_equal(_u32(out[a:b]), orc.count_batch(sq[a:b], se[a:b])), (a, b)


def test_sixty_four_bit_counts_through_the_c_abi(dense):
    from superintervals_b200 import IntervalMap
    s, e, qs, qe, ix, orc = dense
    m = IntervalMap.from_arrays(s, e)
    o = np.argsort(qs, kind="stable")
    assert np.array_equal(m.count_batch_np(qs[o], qe[o]), orc.count_batch(qs[o], qe[o]))


def test_duplicate_coordinates_take_the_second_bitmap_and_the_slow_words():
    """Starts and ends drawn from few distinct coordinates: doubles (d2 bitmap) and triples
    (RB_SLOW words answered from the rank cells)."""
    from superintervals_b200.device import DeviceIndex, ORDER_SORTED
    rng = np.random.default_rng(11)
    n, axis = 200_000, 3_000_000
    s = (rng.integers(0, axis // 3, n) * 3).astype(np.int32)          # ~0.2 values per lattice point: doubles common, triples ~1 %
    e = (s + rng.integers(0, 400, n) * 3).astype(np.int32)
    orc = Oracle(s, e)
    ix = DeviceIndex().build(_dev(s), _dev(e))
    info = ix.bits_info()
    qs = np.sort(rng.integers(-100, axis + 1500, 300_000)).astype(np.int32)
    qe = (qs + rng.integers(0, 900, qs.size)).astype(np.int32)
    got = _u32(ix.count(_dev(qs), _dev(qe), order=ORDER_SORTED))
    assert np.array_equal(got, orc.count_batch(qs, qe)), info
    assert info["built"] and info["slow_words"] > 0
    assert ix.stream_stats()[1] == 0


def test_sparse_index_has_no_rank_bits_and_still_counts():
    from superintervals_b200.device import DeviceIndex, OPT_TIMING, ORDER_SORTED
    rng = np.random.default_rng(13)
    n = 20_000
    s = rng.integers(0, 2_000_000_000, n).astype(np.int32)
    e = (s.astype(np.int64) + rng.integers(0, 100_000, n)).clip(max=2_147_483_000).astype(np.int32)
    ix = DeviceIndex().build(_dev(s), _dev(e))
    assert not ix.bits_info()["built"]
    qs = np.sort(rng.integers(0, 2_000_000_000, 50_000)).astype(np.int32)
    qe = (qs.astype(np.int64) + rng.integers(0, 200_000, qs.size)).clip(max=2_147_483_000).astype(np.int32)
    ix.set_option(OPT_TIMING, 1)
    got = _u32(ix.count(_dev(qs), _dev(qe), order=ORDER_SORTED))
    assert _kernels(ix) == {"count_cells"}
    assert np.array_equal(got, Oracle(s, e).count_batch(qs, qe))


def test_point_index_and_point_queries_stream():
    """Span smaller than one word / stabbing queries / every interval identical."""
    from superintervals_b200.device import DeviceIndex, ORDER_SORTED
    for s, e in ((np.full(70, 5, np.int32), np.full(70, 9, np.int32)),
                 (np.arange(100, dtype=np.int32), np.arange(100, dtype=np.int32)),
                 (np.arange(0, 4000, 2, dtype=np.int32), np.arange(0, 4000, 2, dtype=np.int32) + 37)):
        ix = DeviceIndex().build(_dev(s), _dev(e))
        qs = np.arange(-40, 4100, dtype=np.int32)
        got = _u32(ix.count(_dev(qs), _dev(qs), order=ORDER_SORTED))
        assert np.array_equal(got, Oracle(s, e).count_batch(qs, qs))
