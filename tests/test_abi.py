"""CPU: the C-ABI shared library loads, exports every symbol include/*.h declares, keeps the
reference's struct layouts and host-side bookkeeping, and fails LOUDLY without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from superintervals_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith("#"))   # drop macros
    return set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", txt)) - {"defined", "sizeof"}


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = _lib.lib()
    assert L.si_b200_version().decode() == "0.1.0"


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    declared = (_declared("c_superintervals.h") | _declared("superintervals_b200.h")) - {"size_t", "int32_t"}   # int32_t: return type of the cCombineFn function-pointer typedef
    missing = sorted(s for s in declared if not hasattr(L, s))
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    # and the python binding lists them all
    assert declared <= set(_lib.C_ABI_SYMBOLS + _lib.B200_SYMBOLS), declared - set(_lib.C_ABI_SYMBOLS + _lib.B200_SYMBOLS)


def test_struct_layouts_match_reference():
    # reference c_superintervals.h:61-120 on LP64
    assert C.sizeof(_lib.cSuperIntervals) == 64
    assert C.sizeof(_lib.cIndexResult) == 24 and C.sizeof(_lib.cKeyResult) == 24 and C.sizeof(_lib.cItemResult) == 24
    assert C.sizeof(_lib.Interval) == 12 and C.sizeof(_lib.KeyPair) == 8
    f = {n: getattr(_lib.cSuperIntervals, n).offset for n, _ in _lib.cSuperIntervals._fields_}
    assert f == {"starts": 0, "ends": 8, "data": 16, "branch": 24, "size": 32, "capacity": 40, "idx": 48,
                 "startSorted": 56, "endSorted": 57}


def test_host_side_bookkeeping_without_gpu():
    """create/add/reserve/clear/size + the sortedness flags of addInterval (ref c.h:402-419)."""
    L = _lib.lib()
    si = L.createSuperIntervals()
    c = si.contents
    assert c.size == 0 and c.startSorted and c.endSorted and not c.branch
    L.addInterval(si, 10, 20, 7)
    L.addInterval(si, 10, 15, 8)           # equal start, smaller end: still sorted
    assert c.startSorted and c.endSorted
    L.addInterval(si, 10, 18, 9)           # equal start, larger end -> endSorted drops
    assert c.startSorted and not c.endSorted
    L.addInterval(si, 5, 6, 1)             # smaller start -> startSorted drops
    assert not c.startSorted
    assert L.sizeSuperIntervals(si) == 4 and c.capacity >= 4
    assert [c.starts[i] for i in range(4)] == [10, 10, 10, 5] and [c.data[i] for i in range(4)] == [7, 8, 9, 1]
    iv = _lib.Interval()
    assert L.intervalAt(si, 2, C.byref(iv)) and (iv.start, iv.end, iv.data) == (10, 18, 9)
    assert not L.intervalAt(si, 4, C.byref(iv))
    assert L.startAt(si, 3) == 5 and L.endAt(si, 3) == 6 and L.dataAt(si, 3) == 1
    L.reserveSuperIntervals(si, 100)
    assert c.capacity == 100
    L.clearSuperIntervals(si)
    assert c.size == 0 and c.capacity == 100 and c.startSorted and c.endSorted
    s = np.array([1, 2, 2, 3], np.int32); e = np.array([5, 9, 4, 3], np.int32)
    L.addIntervals(si, s.ctypes.data, e.ctypes.data, None, 4)
    assert c.size == 4 and c.startSorted and c.endSorted and [c.data[i] for i in range(4)] == [0, 1, 2, 3]
    L.destroySuperIntervals(si)
    L.destroySuperIntervals(None)          # no-op, ref c.h:378


def test_result_buffers():
    L = _lib.lib()
    r = L.createIndexResult()
    assert r.size == 0 and r.capacity == 0 and not r.data
    L.clearIndexResult(C.byref(r)); L.destroyIndexResult(C.byref(r))
    k = L.createKeyResult(); L.destroyKeyResult(C.byref(k))
    it = L.createItemResult(); L.destroyItemResult(C.byref(it))


def test_empty_map_queries_are_defined_without_index():
    """ref tests.cpp:223-228: build() on an empty map is a no-op, count == 0, no overlaps."""
    L = _lib.lib()
    L.si_b200_clear_error()
    si = L.createSuperIntervals()
    L.indexSuperIntervals(si)
    assert L.countOverlaps(si, 1, 5) == 0
    assert not L.anyOverlaps(si, 1, 5)
    r = L.createIndexResult()
    L.searchValues(si, 1, 5, C.byref(r))
    assert r.size == 0
    assert L.si_b200_last_error() == 0
    L.destroySuperIntervals(si)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful where no CUDA device exists")
def test_fails_loudly_without_a_gpu():
    """No CPU fallback: indexing without a device latches an error and queries return nothing."""
    L = _lib.lib()
    L.si_b200_clear_error()
    si = L.createSuperIntervals()
    L.addInterval(si, 1, 10, 0)
    L.indexSuperIntervals(si)
    assert L.si_b200_last_error() != 0
    assert b"CUDA" in L.si_b200_last_error_string()
    with pytest.raises(_lib.SuperIntervalsError):
        _lib.check("index")
    assert L.countOverlaps(si, 1, 5) == 0       # not indexed: nothing is computed on the host
    assert L.si_b200_last_error() != 0
    L.si_b200_clear_error()
    L.destroySuperIntervals(si)
    from superintervals_b200 import IntervalMap
    m = IntervalMap()
    m.add(1, 10, "a")
    with pytest.raises(_lib.SuperIntervalsError):
        m.build()
