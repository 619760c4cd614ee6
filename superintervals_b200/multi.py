"""Several GPUs of one node behind the C ABI (include/superintervals_b200.h section 4, csrc/multi.cu):
one host process, one index replica per device, host batches cut into one contiguous range per device,
per-query counts all-gathered over NCCL. Thin ctypes mirror of the siMulti* functions."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class MultiIndex:
    def __init__(self, n_devices: int = 0, devices=None):
        self._L = _lib.lib()
        arr = None
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            n_devices = len(devices)
        self._m = self._L.siMultiCreate(arr, n_devices)
        _lib.check("siMultiCreate")
        if not self._m:
            raise RuntimeError("siMultiCreate failed")

    def __del__(self):
        m, self._m = getattr(self, "_m", None), None
        if m:
            self._L.siMultiDestroy(m)

    @property
    def n_devices(self):
        return int(self._L.siMultiDeviceCount(self._m))

    def index_of(self, rank):
        return self._L.siMultiIndexOf(self._m, rank)

    def build(self, starts, ends, values=None):
        s = np.ascontiguousarray(starts, np.int32)
        e = np.ascontiguousarray(ends, np.int32)
        v = None if values is None else np.ascontiguousarray(values, np.int32)
        if s.shape != e.shape:
            raise ValueError("starts and ends must have the same length")
        self._L.siMultiBuildReplicated(self._m, s.ctypes.data, e.ctypes.data, None if v is None else v.ctypes.data, s.size)
        _lib.check("siMultiBuildReplicated")
        return self

    def count_batch_ptr(self, qs_ptr, qe_ptr, n, out_ptr):
        """Raw-pointer form (pinned host buffers of bench.py)."""
        self._L.siMultiCountBatch(self._m, qs_ptr, qe_ptr, n, out_ptr)
        _lib.check("siMultiCountBatch")

    def count_batch(self, qs, qe):
        qs = np.ascontiguousarray(qs, np.int32)
        qe = np.ascontiguousarray(qe, np.int32)
        if qs.shape != qe.shape:
            raise ValueError("starts and ends must have the same length")
        out = np.zeros(qs.size, np.uint32)
        self.count_batch_ptr(qs.ctypes.data, qe.ctypes.data, qs.size, out.ctypes.data)
        return out

    def search_values_batch_csr(self, qs, qe):
        qs = np.ascontiguousarray(qs, np.int32)
        qe = np.ascontiguousarray(qe, np.int32)
        off = np.zeros(qs.size + 1, np.uint64)
        found = self._L.createIndexResult()
        try:
            self._L.siMultiSearchValuesBatch(self._m, qs.ctypes.data, qe.ctypes.data, qs.size, off.ctypes.data, C.byref(found))
            _lib.check("siMultiSearchValuesBatch")
            n = int(found.size)
            vals = np.ctypeslib.as_array(found.data, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        finally:
            self._L.destroyIndexResult(C.byref(found))
        return off, vals

    def stats(self):
        st = _lib.siMultiStats()
        self._L.siMultiLastStats(self._m, C.byref(st))
        return {k: getattr(st, k) for k, _ in st._fields_}
