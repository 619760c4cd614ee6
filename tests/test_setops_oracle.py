"""CPU: the set-algebra restatement (oracle/si_oracle_setops.c) is pinned against
(1) tests/golden/setops.npz -- outputs of the UNMODIFIED reference C header
    (tools/make_golden_setops.py ran oracle/_ref/libsi_cref.so here), and
(2) the live compiled reference where it is present."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import setops_cases as SC
from oracle.pyoracle import CSetOps, OracleSetOps as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "setops.npz")


def oracle_runner(op, A, B, comb, args):
    if op == "span":
        return O.span(A[0], A[1])
    if op in ("merge", "unique"):
        return getattr(O, op)(*A, combine=comb)
    if op in ("union", "intersection"):
        return getattr(O, op)(A, B, combine=comb)
    if op in ("difference", "symmetric_difference"):
        return getattr(O, op)(A, B)
    return getattr(O, op)(*A, *args)


def check_against_golden(results, with_flags=False):
    g = np.load(GOLD)
    assert len(results) * 1 > 0
    for key, r in results.items():
        if key + "|span" in g.files:
            want = g[key + "|span"]
            assert (r is None and want[0] == 0) or (r is not None and want[0] == 1 and tuple(want[1:]) == tuple(r)), key
            continue
        for name, arr in zip("sed", r[:3]):
            assert np.array_equal(arr, g[key + "|" + name]), (key, name)
        if with_flags:
            assert tuple(bool(x) for x in g[key + "|f"]) == tuple(r[3]), (key, "flags")


def test_oracle_reproduces_the_reference_unit_tests():
    """reference test/tests.cpp:259-373"""
    SC.check_known(oracle_runner)


def test_oracle_matches_reference_fixtures():
    check_against_golden(SC.run_all(oracle_runner))


@pytest.mark.skipif(not CSetOps.reference_available(), reason="compiled reference (oracle/_ref/libsi_cref.so) not present")
def test_oracle_matches_live_reference_on_fresh_seeds():
    ref = CSetOps.reference()
    rng = np.random.default_rng(77)
    for trial in range(60):
        n1, n2 = int(rng.integers(0, 80)), int(rng.integers(0, 80))
        mk = lambda n: (rng.integers(-40, 500, n).astype(np.int32), None, rng.integers(-9, 9, n).astype(np.int32))
        A, B = mk(n1), mk(n2)
        A = (A[0], (A[0] + rng.integers(-3 if trial % 5 == 0 else 0, 45, n1)).astype(np.int32), A[2])
        B = (B[0], (B[0] + rng.integers(0, 70, n2)).astype(np.int32), B[2])
        for op, needs_b, args, combs in SC.OPS:
            for comb in combs:
                got = oracle_runner(op, A, B, comb, args)
                want = ref.run(op, A, B if needs_b else None, combine=comb, args=args)
                if op == "span":
                    assert got == want
                else:
                    assert all(np.array_equal(x, y) for x, y in zip(got, want)), (trial, op, comb, args)
