"""GPU: the Python set-algebra methods of superintervals_b200.IntervalMap against golden vectors
made by the reference's own Python module (tests/golden/py_setops.json, tools/make_golden_py.py).

Results are built maps; among intervals with identical (start, end) the reference's order -- and
with it the order in which payloads are folded into tuples / combined -- is libstdc++'s unstable
std::sort order, so payload folds are compared as multisets of their leaves and
"keep first" results by geometry plus membership."""
import ast
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import py_setops_cases as PC

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "py_setops.json")))


def leaves(rep):
    """repr(value) -> sorted leaves: nested tuples flattened, 'x+y' joins split."""
    v = ast.literal_eval(rep)
    out = []

    def walk(x):
        if isinstance(x, tuple):
            for y in x:
                walk(y)
        elif isinstance(x, str):
            out.extend(x.split("+"))
        else:
            out.append(repr(x))
    walk(v)
    return sorted(out)


def canon(rows, geometry_only):
    if geometry_only:
        return sorted((s, e) for s, e, _ in rows)
    return sorted((s, e, tuple(leaves(r))) for s, e, r in rows)


@pytest.mark.parametrize("name,A,B", PC.cases(), ids=[c[0] for c in PC.cases()])
def test_python_set_algebra_matches_the_reference_module(name, A, B):
    from superintervals_b200 import IntervalMap
    a, b = PC.make(IntervalMap, A), PC.make(IntervalMap, B)
    stored = {(s, e): set() for s, e in zip(A[0], A[1])}
    for s, e, v in zip(*A):
        stored[(s, e)].add(v)
    for key, fn in PC.operations():
        want = GOLD[f"{name}/{key}"]
        got = fn(a, b)
        if key == "span":
            assert (got is None and want is None) or list(got) == list(want), key
            continue
        rows = [[int(s), int(e), repr(v)] for s, e, v in (got.at(i) for i in range(len(got)))]
        first_kept = key in ("merge_first", "unique")
        if key.startswith("intersection"):
            # reference defect (pyx:597-599): intersection() clears self.found_indexes but appends to
            # other.found_indexes, so stale hits of earlier intervals are re-tested and every true pair is
            # emitted 1 + (earlier intervals sharing that partner) times. C++ (hpp:1167) and C (c.h:927)
            # clear per interval. We return each pair once: equal as sets, never more than the reference.
            assert set(canon(rows, False)) == set(canon(want, False)), (name, key)
            assert len(rows) == len(set(canon(rows, False))) <= len(want), (name, key)
            continue
        assert canon(rows, first_kept) == canon(want, first_kept), (name, key)
        if key == "unique":
            assert all(ast.literal_eval(r) in stored[(s, e)] for s, e, r in rows)
        # a result is a built map: it answers queries
        if len(got):
            s0, e0, _ = got.at(0)
            assert got.count(s0, e0) >= 1 and got.has_overlaps(s0, e0)
