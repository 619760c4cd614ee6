"""Shared helpers for the test-suite (not collected)."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
NONE64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def golden_files(tag=None):
    pat = f"*.{tag}.npz" if tag else "*.npz"
    return sorted(glob.glob(os.path.join(GOLDEN, pat)))


def load_vectors():
    with open(os.path.join(GOLDEN, "tests_cpp_vectors.json")) as f:
        return json.load(f)["cases"]


def brute_force(starts, ends, qs, qe):
    """O(N*Q) truth on the POSITION-SORTED arrays: hit set {j : starts[j] <= qe and ends[j] >= qs}."""
    s = np.asarray(starts, np.int64)[None, :]
    e = np.asarray(ends, np.int64)[None, :]
    hit = (s <= np.asarray(qe, np.int64)[:, None]) & (e >= np.asarray(qs, np.int64)[:, None])
    return hit


def csr_from_hits(hit):
    """Descending-index CSR (offsets, idx) from a boolean (Q, N) matrix."""
    counts = hit.sum(1)
    off = np.zeros(len(counts) + 1, np.uint64)
    np.cumsum(counts, out=off[1:])
    idx = np.concatenate([np.flatnonzero(r)[::-1] for r in hit]) if len(hit) else np.zeros(0, np.int64)
    return off, idx.astype(np.uint32)


def canonical_values(offsets, values, keys):
    """Order-insensitive view of a CSR value list for inputs whose exact-duplicate (start,end)
    groups have an implementation-defined order: sort values inside runs of
    identical keys of each query's list."""
    out = values.copy()
    for q in range(len(offsets) - 1):
        a, b = int(offsets[q]), int(offsets[q + 1])
        i = a
        while i < b:
            j = i + 1
            while j < b and keys[j, 0] == keys[i, 0] and keys[j, 1] == keys[i, 1]:
                j += 1
            if j - i > 1:
                out[i:j] = np.sort(out[i:j])
            i = j
    return out
