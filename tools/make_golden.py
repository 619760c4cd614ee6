"""Generate tests/golden/*.npz by running the REAL reference (oracle/_ref, compiled in place
from /root/reference) on small seeded inputs. Run in the build container only:

    make -C oracle && python tools/make_golden.py

Each fixture stores the inputs and the reference's own outputs:
  index arrays after build() (starts, ends, data, branch), count / count_linear /
  count_large, has_overlaps, search_values CSR (+ search_values_large), search_keys,
  search_idxs (C++ vector overload: first run ascending), coverage.
`presorted` fixtures feed the reference intervals already in (start asc, end desc,
insertion) order, so it performs no sort (hpp:1416,1421) and its payload order is
fully determined; `shuffled` fixtures carry the unstable std::sort tie order of the
reference and are compared on order-insensitive outputs only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference  # noqa: E402
from superintervals_b200 import workloads as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def canonical_order(s, e):
    # (start asc, end desc, insertion idx): np.lexsort is stable, last key is primary
    return np.lexsort((-(e.astype(np.int64)), s.astype(np.int64)))


def cases():
    rng = np.random.default_rng(2024)
    yield "c1_small", W.config1(3000, 0)
    s, e, qs, qe = W.config2(3000, 1500, 2, axis=600_000)
    yield "c2_dense", (s, e, qs, qe)
    yield "c3_nested", W.config3(4000, 3000, 42, axis=400_000)
    n, nq = 800, 600
    s = rng.integers(0, 60, n).astype(np.int32)
    e = (s + rng.integers(0, 9, n)).astype(np.int32)
    qs = rng.integers(-5, 70, nq).astype(np.int32)
    qe = (qs + rng.integers(-2, 12, nq)).astype(np.int32)      # includes inverted queries (Q6)
    yield "dups_inverted", (s, e, qs, qe)
    s = np.arange(2000, dtype=np.int32)
    e = (s + 1 + (np.arange(2000) % 5)).astype(np.int32)
    s[0], e[0] = 0, 100_000                                     # giant container over a staircase
    qs = rng.integers(0, 2100, 1000).astype(np.int32)
    qe = (qs + rng.integers(0, 60, 1000)).astype(np.int32)
    yield "container", (s, e, qs, qe)
    s = rng.integers(-2_000_000_000, 2_000_000_000, 2000).astype(np.int64)
    e = np.minimum(s + rng.integers(0, 400_000_000, 2000), 2_147_483_000).astype(np.int32)
    qs = rng.integers(-2_100_000_000, 2_100_000_000, 500).astype(np.int64)
    qe = np.clip(qs + rng.integers(0, 500_000_000, 500), -2_147_483_648, 2_147_483_647).astype(np.int32)
    yield "signed_wide", (s.astype(np.int32), e, qs.astype(np.int32), qe)


def run(name, s, e, qs, qe, presorted):
    if presorted:
        o = canonical_order(s, e)
        s, e = np.ascontiguousarray(s[o]), np.ascontiguousarray(e[o])
    ref = Reference(s, e)                      # data = insertion index, as test/bench.cpp:210
    rs, re_, rd, rb = ref.export()
    out = dict(in_starts=s, in_ends=e, qs=qs, qe=qe, flags=np.int32(ref.flags()),
               starts=rs, ends=re_, data=rd, branch=rb,
               count=ref.count_batch(qs, qe, 0), count_linear=ref.count_batch(qs, qe, 1),
               count_large=ref.count_batch(qs, qe, 2), has_overlaps=ref.has_overlaps_batch(qs, qe))
    off, vals = ref.search_values_batch(qs, qe, 0)
    off_l, vals_l = ref.search_values_batch(qs, qe, 1)
    assert np.array_equal(off, off_l)
    koff, keys = ref.search_keys_batch(qs, qe)
    ioff, idxs = ref.search_idxs_batch(qs, qe)
    cc, cv = ref.coverage_batch(qs, qe)
    out.update(offsets=off, values=vals, values_large=vals_l, keys=keys, idxs_cpp=idxs.astype(np.uint32),
               cov_count=cc, cov_sum=cv)
    tag = "presorted" if presorted else "shuffled"
    np.savez_compressed(os.path.join(OUT, f"{name}.{tag}.npz"), **out)
    print(f"{name}.{tag}: n={len(s)} nq={len(qs)} hits={int(off[-1])}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    print("reference build:", Reference.lib() and Reference.kind)
    for name, (s, e, qs, qe) in cases():
        run(name, s, e, qs, qe, True)
        run(name, s, e, qs, qe, False)
