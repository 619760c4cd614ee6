"""Seeded input pairs for the set-algebra tests (shared by the CPU oracle tests, the GPU parity
tests and tools/make_golden_setops.py). Each case is (name, A, B) with A, B = (starts, ends, data)
int32 arrays in STORED order."""
import numpy as np

I32 = np.iinfo(np.int32)


def _rand(rng, n, span, maxlen, sort=False, data_span=50):
    s = rng.integers(0, span, n).astype(np.int32)
    e = (s + rng.integers(0, maxlen, n)).astype(np.int32)
    d = rng.integers(-data_span, data_span, n).astype(np.int32)
    if sort:
        o = np.lexsort((-e, s))
        s, e, d = s[o], e[o], d[o]
    return s, e, d


def cases(scale=1):
    rng = np.random.default_rng(2024)
    out = []
    n = 300 * scale
    out.append(("random_shuffled", _rand(rng, n, 40 * n, 120), _rand(rng, n, 40 * n, 150)))
    out.append(("random_presorted", _rand(rng, n, 40 * n, 120, sort=True), _rand(rng, n, 40 * n, 150, sort=True)))
    dspan = max(60, n // 5)                                                                      # ~5 intervals per start coordinate
    out.append(("dense_dups", _rand(rng, n, dspan, 9), _rand(rng, n, dspan, 7)))                # many exact duplicates, shared endpoints
    out.append(("sparse", _rand(rng, n // 3, 4_000 * n, 30), _rand(rng, n // 3, 4_000 * n, 30)))   # almost no overlaps
    a = _rand(rng, n // 2, 30 * n, 80)
    big = (np.array([5, 10 * n, 3], np.int32), np.array([25 * n, 28 * n, 2 * n], np.int32), np.array([7, 8, 9], np.int32))
    out.append(("containers", tuple(np.concatenate([x, y]) for x, y in zip(a, big)), _rand(rng, n // 2, 30 * n, 400)))
    s = rng.integers(-2_000_000_000, 2_000_000_000, n, dtype=np.int64)
    e = np.minimum(s + rng.integers(0, 60_000_000 // scale, n), I32.max - 1)
    out.append(("signed_wide", (s.astype(np.int32), e.astype(np.int32), rng.integers(-9, 9, n).astype(np.int32)),
                _rand(rng, n, 1_000_000 if scale == 1 else 20_000 * n, 5_000)))
    out.append(("empty_a", (np.zeros(0, np.int32),) * 3, _rand(rng, 20, 500, 30)))
    out.append(("empty_b", _rand(rng, 20, 500, 30), (np.zeros(0, np.int32),) * 3))
    out.append(("single", (np.array([7], np.int32), np.array([19], np.int32), np.array([1], np.int32)),
                (np.array([10], np.int32), np.array([12], np.int32), np.array([2], np.int32))))
    out.append(("adjacent", (np.array([1, 6, 11, 11], np.int32), np.array([5, 10, 15, 11], np.int32), np.array([1, 2, 3, 4], np.int32)),
                (np.array([5, 16], np.int32), np.array([6, 20], np.int32), np.array([9, 8], np.int32))))
    # malformed stored intervals (start > end) in A and in B: the exact-order paths
    a = _rand(rng, 150, 3_000, 60)
    bad = rng.random(150) < 0.15
    a = (a[0], np.where(bad, a[0] - rng.integers(1, 40, 150), a[1]).astype(np.int32), a[2])
    b = _rand(rng, 150, 3_000, 60)
    bad = rng.random(150) < 0.15
    b = (b[0], np.where(bad, b[0] - rng.integers(1, 40, 150), b[1]).astype(np.int32), b[2])
    out.append(("malformed", a, b))
    return out


# (op, needs B, extra args, combiner names to try)
OPS = [("merge", False, (), (None, "sum", "second")), ("unique", False, (), (None, "sum")),
       ("union", True, (), (None, "sum")), ("intersection", True, (), (None, "sum", "second")),
       ("difference", True, (), (None,)), ("symmetric_difference", True, (), (None,)),
       ("gaps", False, (-7, 9_000, 42), (None,)), ("gaps", False, (100, 50, 1), (None,)),
       ("expand", False, (5, 11, -100, 11_000), (None,)), ("expand", False, (-30, -25, I32.min, I32.max), (None,)),
       ("flank", False, (6, 9, 3, 10_500), (None,)), ("flank", False, (0, 4, I32.min, I32.max), (None,)),
       ("span", False, (), (None,))]


def run_all(runner, scale=1, flags=False):
    """runner(op, A, B, combine, args) -> result; returns {key: result} over every case x op."""
    res = {}
    for name, A, B in cases(scale):
        for op, needs_b, args, combs in OPS:
            for comb in combs:
                res[f"{name}/{op}/{comb}/{args}"] = runner(op, A, B if needs_b else None, comb, args)
    return res


# Known answers from the reference's own unit tests (reference test/tests.cpp:259-373):
# (op, A, B, combine, args, expected [(start, end), ...] or span tuple, expected data or None)
def _t(*rows):
    return tuple(np.array(c, np.int32) for c in zip(*rows)) if rows else (np.zeros(0, np.int32),) * 3


KNOWN = [
    ("merge", _t((1, 5, 0), (3, 8, 1), (20, 30, 2)), None, None, (), [(1, 8), (20, 30)], [0, 2]),            # tests.cpp:259-272
    ("merge", _t((1, 5, 0), (3, 8, 1), (20, 30, 2)), None, "sum", (), [(1, 8), (20, 30)], [1, 2]),          # tests.cpp:274-277
    ("gaps", _t((10, 20, 0), (30, 40, 1)), None, None, (0, 50, 0), [(0, 9), (21, 29), (41, 50)], None),       # tests.cpp:280-293
    ("union", _t((1, 10, 0)), _t((5, 25, 1)), None, (), [(1, 25)], [0]),                                      # tests.cpp:295-306
    ("intersection", _t((1, 10, 0), (20, 30, 1)), _t((5, 25, 2)), None, (), [(5, 10), (20, 25)], [0, 1]),    # tests.cpp:308-322
    ("difference", _t((1, 10, 0)), _t((4, 6, 1)), None, (), [(1, 3), (7, 10)], [0, 0]),                       # tests.cpp:331-343
    ("symmetric_difference", _t((1, 10, 0)), _t((5, 15, 1)), None, (), [(1, 4), (11, 15)], [0, 1]),          # tests.cpp:345-357
    ("span", _t((10, 20, 0), (5, 8, 1), (15, 50, 2)), None, None, (), (5, 50), None),                          # tests.cpp:359-369
    ("span", _t(), None, None, (), None, None),                                                               # tests.cpp:371-372
]


def check_known(runner):
    for op, A, B, comb, args, geom, data in KNOWN:
        r = runner(op, A, B, comb, args)
        if op == "span":
            assert r == geom, (op, r)
            continue
        assert list(zip(r[0].tolist(), r[1].tolist())) == geom, (op, r)
        if data is not None:
            assert r[2].tolist() == data, (op, r)
