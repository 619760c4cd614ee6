"""Python ``IntervalMap`` over libsuperintervals_b200.so.

Host-side mirror of the reference's Python class for the batch overlap-query
path (reference src/superintervals/intervalmap.pyx:18-494): same method names,
argument meaning and error behaviour (ValueError on length mismatch,
IndexError on out-of-range ``at``). Every query executes on the GPU through the
C ABI (include/c_superintervals.h, include/superintervals_b200.h); payload
objects stay on the host and are addressed by the int32 insertion index the
device index carries.

Additions to the reference API (the CSR forms the GPU path produces natively):
``count_batch_np``, ``search_values_batch_csr``, ``search_idxs_batch_csr``,
``search_keys_batch_csr``, ``has_overlaps_batch``, ``coverage_batch``.
The ``*_batch`` methods of the reference return Python lists / list-of-lists
(pyx:395-400, 436-446, 482-494); they are kept and built from the CSR forms.

Set algebra (pyx:510-819: merge_overlaps, union_with, intersection, difference,
symmetric_difference, gaps, span, expand, flank, unique): the geometry comes from the
library's device implementation of the reference's C functions
(include/c_superintervals.h "set operations"); Python payload objects are folded on the
host exactly as the reference does (tuples by default, or the caller's ``combine``). As
in the reference every result is a new, BUILT map.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["IntervalMap"]


def _as_i32(a, name):
    arr = np.ascontiguousarray(a, dtype=np.int32)
    if arr.ndim != 1:
        raise ValueError(f"{name} must be one-dimensional")
    return arr


def _vector_overload_order(offsets, idx, ub):
    """Reorder all-descending CSR hit lists (the walk's order: C ABI, device kernels) into the order of the
    reference's C++ `search_idxs(start, end, vector&)` (hpp:879-905), which the reference's Python class uses:
    the first contiguous run of positions (ub, ub-1, ... down to the first miss; ub[q] = upper_bound(end) of query q)
    comes out ASCENDING, the hits the branch walk finds after it stay descending. A list that does not begin at ub
    has an empty first run. Vectorised over the whole batch."""
    idx = np.asarray(idx)
    total = idx.shape[0]
    if total == 0:
        return idx
    off = np.asarray(offsets, np.int64)
    lens = np.diff(off)
    nz = lens > 0
    lo = np.repeat(off[:-1][nz], lens[nz])                 # segment start of every hit
    pos = np.arange(total, dtype=np.int64) - lo             # rank of the hit inside its list
    in_run = idx.astype(np.int64) == np.repeat(np.asarray(ub, np.int64)[nz], lens[nz]) - pos
    # run length = first position of the list where the descent by one breaks
    brk = np.where(in_run, np.int64(np.iinfo(np.int64).max), pos)
    run = np.minimum(np.minimum.reduceat(brk, off[:-1][nz]), lens[nz])
    run = np.repeat(run, lens[nz])
    src = np.where(pos < run, lo + run - 1 - pos, lo + pos)
    return idx[src]


class IntervalMap:
    """SuperIntervals interval map: end-inclusive intervals with associated Python objects."""

    def __init__(self):
        self._L = _lib.lib()
        self._si = self._L.createSuperIntervals()
        if not self._si:
            raise MemoryError("createSuperIntervals failed")
        self._values = []          # payload objects in insertion order
        self._built = False

    def __del__(self):
        si, self._si = getattr(self, "_si", None), None
        if si:
            self._L.destroySuperIntervals(si)

    # ---- container protocol (pyx:38-42) --------------------------------------------------
    def __len__(self):
        return self.size()

    def __getitem__(self, index):
        return self.at(index)

    # ---- building (pyx:44-137) -------------------------------------------------------------
    def add(self, start, end, value=None):
        """Add an interval [start, end] (inclusive) with an associated Python object."""
        self._L.addInterval(self._si, int(start), int(end), len(self._values))
        self._values.append(value)
        self._built = False

    @classmethod
    def from_arrays(cls, starts, ends, values=None):
        """Create a ready-to-query IntervalMap from arrays (pyx:63-131)."""
        self = cls()
        s = _as_i32(starts, "starts")
        e = _as_i32(ends, "ends")
        if s.shape[0] != e.shape[0]:
            raise ValueError("starts and ends must have the same length")
        n = s.shape[0]
        if values is not None and len(values) != n:
            raise ValueError("values length must match starts/ends length")
        self._L.addIntervals(self._si, s.ctypes.data, e.ctypes.data, None, n)
        self._values = list(values) if values is not None else [None] * n
        self.build()
        return self

    def build(self):
        """Build the superintervals index (device sort + branch pass). Required before queries."""
        self._L.indexSuperIntervals(self._si)
        _lib.check("build")
        self._built = True

    def clear(self):
        self._L.clearSuperIntervals(self._si)
        self._values = []
        self._built = False

    def reserve(self, n):
        self._L.reserveSuperIntervals(self._si, int(n))

    def size(self):
        return int(self._L.sizeSuperIntervals(self._si))

    # ---- element access (pyx:139-211) -------------------------------------------------------
    def _check_index(self, index):
        if self.size() == 0 or index < 0 or index >= self.size():
            raise IndexError("Index out of range")

    def at(self, index):
        self._check_index(index)
        c = self._si.contents
        return c.starts[index], c.ends[index], self._values[c.data[index]]

    def starts_at(self, index):
        self._check_index(index)
        return self._si.contents.starts[index]

    def ends_at(self, index):
        self._check_index(index)
        return self._si.contents.ends[index]

    def data_at(self, index):
        self._check_index(index)
        return self._values[self._si.contents.data[index]]

    # index arrays in position order (host mirrors of the device index)
    def _mirror(self, field, dtype):
        n = self.size()
        if n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(getattr(self._si.contents, field), shape=(n,)).astype(dtype)

    @property
    def starts(self): return self._mirror("starts", np.int32)
    @property
    def ends(self): return self._mirror("ends", np.int32)
    @property
    def data_index(self): return self._mirror("data", np.int32)
    @property
    def branch(self):
        if not self._si.contents.branch:
            return np.zeros(0, np.uint64)
        return self._mirror("branch", np.uint64)

    # ---- single queries (pyx:241-361) ---------------------------------------------------------
    def has_overlaps(self, start, end):
        r = bool(self._L.anyOverlaps(self._si, int(start), int(end)))
        _lib.check("has_overlaps")
        return r

    def count(self, start, end):
        r = int(self._L.countOverlaps(self._si, int(start), int(end)))
        _lib.check("count")
        return r

    def _one(self, start, end):
        return (np.array([start], np.int32), np.array([end], np.int32))

    def search_values(self, start, end):
        _, vals = self.search_values_batch_csr(*self._one(start, end))
        return [self._values[i] for i in vals]

    def _idxs_in_vector_order(self, starts, ends):
        off, idx = self.search_idxs_batch_csr(starts, ends)
        if idx.size == 0:
            return off, idx
        ub = np.searchsorted(self._mirror("starts", np.int32), np.asarray(ends, np.int32), "right") - 1   # hpp:501-516
        return off, _vector_overload_order(off, idx, ub)

    def search_idxs(self, start, end):
        """Positions of overlapping intervals in the order of the reference's Python class: it calls the C++
        vector overload (pyx:299 -> hpp:879-905), which emits the first contiguous run ASCENDING and the rest
        of the walk descending. The C ABI and the CSR forms return all-descending lists."""
        return [int(i) for i in self._idxs_in_vector_order(*self._one(start, end))[1]]

    def search_keys(self, start, end):
        # pyx:314-320: search_idxs, then (starts[i], ends[i]) per position -- same ordering quirk
        _, idx = self._idxs_in_vector_order(*self._one(start, end))
        c = self._si.contents
        return [(int(c.starts[i]), int(c.ends[i])) for i in idx]

    def search_items(self, start, end):
        # pyx:335-345: search_idxs, then (start, end, data) per position -- same ordering quirk
        _, idx = self._idxs_in_vector_order(*self._one(start, end))
        c = self._si.contents
        return [(int(c.starts[i]), int(c.ends[i]), self._values[c.data[i]]) for i in idx]

    def coverage(self, start, end):
        cnt, cov = C.c_size_t(0), C.c_int32(0)
        self._L.coverage(self._si, int(start), int(end), C.byref(cnt), C.byref(cov))
        _lib.check("coverage")
        return int(cnt.value), int(cov.value)

    # ---- batch queries: numpy / CSR forms -------------------------------------------------------
    def _pair(self, starts, ends):
        s = _as_i32(starts, "starts")
        e = _as_i32(ends, "ends")
        if s.shape[0] != e.shape[0]:
            raise ValueError("starts and ends must have the same length")   # pyx:392-393
        return s, e

    def count_batch_np(self, starts, ends):
        s, e = self._pair(starts, ends)
        out = np.zeros(s.shape[0], np.uint64)
        self._L.countOverlapsBatch(self._si, s.ctypes.data, e.ctypes.data, s.shape[0], out.ctypes.data)
        _lib.check("count_batch")
        return out

    def has_overlaps_batch(self, starts, ends):
        s, e = self._pair(starts, ends)
        out = np.zeros(s.shape[0], np.bool_)
        self._L.anyOverlapsBatch(self._si, s.ctypes.data, e.ctypes.data, s.shape[0], out.ctypes.data)
        _lib.check("has_overlaps_batch")
        return out

    def coverage_batch(self, starts, ends):
        s, e = self._pair(starts, ends)
        cnt = np.zeros(s.shape[0], np.uint64)
        cov = np.zeros(s.shape[0], np.int32)
        self._L.coverageBatch(self._si, s.ctypes.data, e.ctypes.data, s.shape[0], cnt.ctypes.data, cov.ctypes.data)
        _lib.check("coverage_batch")
        return cnt, cov

    def _csr(self, fn, make, destroy, starts, ends, shape_tail, dtype):
        s, e = self._pair(starts, ends)
        n = s.shape[0]
        offsets = np.zeros(n + 1, np.uint64)
        found = make()
        fn(self._si, s.ctypes.data, e.ctypes.data, n, offsets.ctypes.data, C.byref(found))
        _lib.check(fn.__name__)
        total = int(found.size)
        if total:
            buf = np.ctypeslib.as_array(C.cast(found.data, C.POINTER(C.c_int32)), shape=(total,) + shape_tail)
            out = buf.astype(dtype, copy=True)
        else:
            out = np.zeros((0,) + shape_tail, dtype)
        destroy(C.byref(found))
        return offsets, out

    def search_values_batch_csr(self, starts, ends):
        """(offsets[n+1], idx[total]): idx are insertion indices into the payload list."""
        return self._csr(self._L.searchValuesBatch, self._L.createIndexResult, self._L.destroyIndexResult,
                         starts, ends, (), np.int32)

    def search_idxs_batch_csr(self, starts, ends):
        return self._csr(self._L.searchIdxsBatch, self._L.createIndexResult, self._L.destroyIndexResult,
                         starts, ends, (), np.int32)

    def search_keys_batch_csr(self, starts, ends):
        return self._csr(self._L.searchKeysBatch, self._L.createKeyResult, self._L.destroyKeyResult,
                         starts, ends, (2,), np.int32)

    # ---- batch queries: the reference's list forms (pyx:363-494) -----------------------------------
    def count_batch(self, starts, ends):
        return [int(c) for c in self.count_batch_np(starts, ends)]

    def search_idxs_batch(self, starts, ends):
        off, idx = self._idxs_in_vector_order(starts, ends)          # pyx:440: the C++ vector overload per query
        return [[int(i) for i in idx[int(off[q]):int(off[q + 1])]] for q in range(len(off) - 1)]

    def search_values_batch(self, starts, ends):
        off, idx = self.search_values_batch_csr(starts, ends)
        vals = self._values
        return [[vals[i] for i in idx[int(off[q]):int(off[q + 1])]] for q in range(len(off) - 1)]

    # ---- set algebra (pyx:510-819) ------------------------------------------------------------------
    def _stored(self):
        """Stored arrays (starts, ends, payload index) in the handle's current order."""
        return self._mirror("starts", np.int32), self._mirror("ends", np.int32), self._mirror("data", np.int32)

    @classmethod
    def _wrap(cls, L, si, values):
        """A map over a handle the library returned; `values[i]` is the payload of its i-th interval."""
        m = cls.__new__(cls)
        m._L, m._si, m._values, m._built = L, si, list(values), False
        n = int(si.contents.size)
        if n:
            np.ctypeslib.as_array(si.contents.data, shape=(n,))[:] = np.arange(n, dtype=np.int32)
        m.build()
        return m

    def _take(self, si):
        n = int(si.contents.size)
        if n == 0:
            z = np.zeros(0, np.int32)
            return z, z, z
        return tuple(np.ctypeslib.as_array(getattr(si.contents, f), shape=(n,)).copy() for f in ("starts", "ends", "data"))

    @staticmethod
    def _fold(acc, d, combine):
        if combine is not None:
            return combine(acc, d)
        return acc + (d,) if isinstance(acc, tuple) else (acc, d)          # pyx:549-550

    def _merged(self, si, starts, ends, vals, combine):
        """Fold payloads into the clusters of a merged handle: members in (start, end) order (pyx:522)."""
        ms, _, _ = self._take(si)
        out = [None] * ms.size
        seen = np.zeros(ms.size, bool)
        order = np.lexsort((ends, starts))
        cl = np.searchsorted(ms, starts[order], "right") - 1
        for k, c in zip(order, cl):
            if seen[c]:
                out[c] = self._fold(out[c], vals[k], combine)
            else:
                out[c], seen[c] = vals[k], True
        return self._wrap(self._L, si, out)

    def merge_overlaps(self, combine=None):
        """Coalesce overlapping intervals into a disjoint set (pyx:525-556); merged payloads are
        collected into a tuple unless `combine(a, b)` is given."""
        s, e, d = self._stored()
        si = self._L.mergeOverlaps(self._si, None)
        _lib.check("merge_overlaps")
        return self._merged(si, s, e, [self._values[i] for i in d], combine)

    def union_with(self, other, combine=None):
        """All covered regions of both maps, coalesced (pyx:558-575)."""
        s1, e1, d1 = self._stored()
        s2, e2, d2 = other._stored()
        si = self._L.unionWith(self._si, other._si, None)
        _lib.check("union_with")
        vals = [self._values[i] for i in d1] + [other._values[i] for i in d2]
        return self._merged(si, np.concatenate([s1, s2]), np.concatenate([e1, e2]), vals, combine)

    def intersection(self, other, combine=None):
        """Every overlapping sub-region of the two maps, not coalesced (pyx:577-609); payloads are
        paired into (a, b) unless `combine(a, b)` is given. `other` must be built. Every overlapping
        pair is returned once, as the C++ and C implementations do (hpp:1164-1179, c.h:921-939); the
        reference's Python method re-emits pairs because it never clears `other.found_indexes`
        (pyx:597-599) -- its output is this one with duplicates."""
        b = self._L.createIndexResult()
        si = self._L.intersectionPairs(self._si, other._si, C.byref(b))
        _lib.check("intersection")
        _, _, da = self._take(si)
        db = [b.data[i] for i in range(int(b.size))]
        self._L.destroyIndexResult(C.byref(b))
        pair = combine if combine is not None else (lambda x, y: (x, y))
        return self._wrap(self._L, si, [pair(self._values[i], other._values[j]) for i, j in zip(da, db)])

    def difference(self, other):
        """Regions of this map not covered by `other` (pyx:611-649); pieces inherit their source payload."""
        si = self._L.difference(self._si, other._si)
        _lib.check("difference")
        _, _, d = self._take(si)
        return self._wrap(self._L, si, [self._values[i] for i in d])

    def symmetric_difference(self, other):
        """Regions in exactly one of the two maps (pyx:651-664). Both must be built."""
        return self.difference(other).union_with(other.difference(self))

    def gaps(self, lo, hi, fill=None):
        """The uncovered regions inside [lo, hi] (pyx:666-695)."""
        si = self._L.intervalGaps(self._si, int(lo), int(hi), 0)
        _lib.check("gaps")
        return self._wrap(self._L, si, [fill] * int(si.contents.size))

    def span(self):
        """(min start, max end), or None when empty (pyx:697-715)."""
        lo, hi = C.c_int32(0), C.c_int32(0)
        ok = self._L.intervalSpan(self._si, C.byref(lo), C.byref(hi))
        _lib.check("span")
        return (lo.value, hi.value) if ok else None

    _I32 = np.iinfo(np.int32)

    def expand(self, left, right, lo=None, hi=None):
        """Grow or shrink every interval, like `bedtools slop` (pyx:717-746)."""
        si = self._L.expandIntervals(self._si, int(left), int(right), self._I32.min if lo is None else int(lo),
                                     self._I32.max if hi is None else int(hi))
        _lib.check("expand")
        _, _, d = self._take(si)
        return self._wrap(self._L, si, [self._values[i] for i in d])

    def flank(self, left, right, lo=None, hi=None):
        """The strips beside each interval, like `bedtools flank` (pyx:748-788); originals are not emitted."""
        si = self._L.flankIntervals(self._si, int(left), int(right), self._I32.min if lo is None else int(lo),
                                    self._I32.max if hi is None else int(hi))
        _lib.check("flank")
        _, _, d = self._take(si)
        return self._wrap(self._L, si, [self._values[i] for i in d])

    def unique(self, combine=None):
        """One interval per distinct (start, end) pair (pyx:790-819); the first payload is kept
        unless `combine(a, b)` folds the duplicates."""
        s, e, d = self._stored()
        si = self._L.uniqueIntervals(self._si, None)
        _lib.check("unique")
        us, ue, ud = self._take(si)
        if combine is None:
            return self._wrap(self._L, si, [self._values[i] for i in ud])
        acc = {}
        for k in np.lexsort((e, s)):
            key = (int(s[k]), int(e[k]))
            acc[key] = combine(acc[key], self._values[d[k]]) if key in acc else self._values[d[k]]
        return self._wrap(self._L, si, [acc[(int(a), int(b))] for a, b in zip(us, ue)])
