"""Inputs and the operation list shared by tools/make_golden_py.py (run on the reference's Python
module) and tests/test_gpu_python_setops.py (run on superintervals_b200.IntervalMap)."""
import numpy as np


def cases():
    rng = np.random.default_rng(31)
    out = []
    for name, n, span, mx in (("small", 40, 300, 40), ("dups", 60, 30, 6), ("sparse", 50, 5000, 25)):
        sets = []
        for tag in "ab":
            s = rng.integers(0, span, n)
            e = s + rng.integers(0, mx, n)
            sets.append(([int(x) for x in s], [int(x) for x in e], [f"{tag}{i}" for i in range(n)]))
        out.append((name, sets[0], sets[1]))
    out.append(("empty", ([], [], []), ([3], [9], ["b0"])))
    return out


def make(cls, S):
    m = cls()
    for s, e, v in zip(*S):
        m.add(s, e, v)
    m.build()
    return m


def operations():
    first = lambda x, y: x
    join = lambda x, y: f"{x}+{y}"
    return [("merge", lambda a, b: a.merge_overlaps()), ("merge_first", lambda a, b: a.merge_overlaps(first)),
            ("merge_join", lambda a, b: a.merge_overlaps(join)),
            ("union", lambda a, b: a.union_with(b)), ("union_join", lambda a, b: a.union_with(b, join)),
            ("intersection", lambda a, b: a.intersection(b)), ("intersection_join", lambda a, b: a.intersection(b, join)),
            ("difference", lambda a, b: a.difference(b)), ("symmetric_difference", lambda a, b: a.symmetric_difference(b)),
            ("gaps", lambda a, b: a.gaps(-5, 350, "gap")), ("span", lambda a, b: a.span()),
            ("expand", lambda a, b: a.expand(3, 7)), ("expand_clamped", lambda a, b: a.expand(10, -4, 5, 280)),
            ("flank", lambda a, b: a.flank(4, 6)), ("flank_clamped", lambda a, b: a.flank(9, 2, 2, 290)),
            ("unique", lambda a, b: a.unique()), ("unique_join", lambda a, b: a.unique(join))]
