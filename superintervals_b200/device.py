"""Device-resident front-end: queries and results stay in HBM.

PyTorch is plumbing only here -- it owns device memory and the CUDA stream;
every kernel is launched by libsuperintervals_b200.so through the raw-pointer
entry points of include/superintervals_b200.h (section 3). Tensors are passed
as ``data_ptr()`` and the current torch stream as ``cuda_stream``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (OPT_STREAM, OPT_STREAM_BUDGET, OPT_L2_PERSIST, OPT_NARROW_SORT, OPT_STAB_BUDGET, OPT_STAB_LISTS, COUNT_AUTO, COUNT_CELLS, COUNT_RANK, COUNT_WALK, OPT_CELLS_DIRECT_BYTES, OPT_CELLS_FILL, FILL_IDXS, FILL_ITEMS, FILL_KEYS, FILL_VALUES, OPT_BUCKET_INTERVALS,
                   OPT_COUNT_ALGO, OPT_TIMING, OPT_WINDOW_SHIFT, ORDER_ASIS, ORDER_AUTO, ORDER_SORTED, ORDER_UNSORTED)

__all__ = ["DeviceIndex", "ORDER_AUTO", "ORDER_SORTED", "ORDER_UNSORTED", "ORDER_ASIS", "OPT_COUNT_ALGO",
           "OPT_BUCKET_INTERVALS", "OPT_WINDOW_SHIFT", "OPT_TIMING", "COUNT_AUTO", "COUNT_WALK", "COUNT_RANK", "COUNT_CELLS",
           "OPT_CELLS_DIRECT_BYTES", "OPT_CELLS_FILL", "OPT_STAB_LISTS", "OPT_STAB_BUDGET", "OPT_STREAM", "OPT_STREAM_BUDGET", "OPT_L2_PERSIST", "OPT_NARROW_SORT"]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_i32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int32 and t.is_contiguous() and t.dim() == 1):
        raise ValueError(f"{name} must be a contiguous 1-D int32 CUDA tensor")
    return t


class DeviceIndex:
    """A superset index living on one GPU (the current torch device at construction)."""

    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError("superintervals_b200 needs a CUDA device; there is no CPU fallback")
        self._L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ix = self._L.siIndexCreate()
        _lib.check("siIndexCreate")
        if not self._ix:
            raise RuntimeError("siIndexCreate failed")

    def __del__(self):
        ix, self._ix = getattr(self, "_ix", None), None
        if ix:
            self._L.siIndexDestroy(ix)

    def __len__(self):
        return int(self._L.siIndexSize(self._ix))

    @property
    def device_bytes(self):
        return int(self._L.siIndexDeviceBytes(self._ix))

    # ---- build ------------------------------------------------------------------------------
    def build(self, starts, ends, values=None):
        """build() from device tensors (int32). values=None stores the insertion index."""
        _chk_i32(starts, "starts"); _chk_i32(ends, "ends")
        if starts.numel() != ends.numel():
            raise ValueError("starts and ends must have the same length")
        if values is not None:
            _chk_i32(values, "values")
        rc = self._L.siIndexBuildDevice(self._ix, starts.data_ptr(), ends.data_ptr(),
                                        values.data_ptr() if values is not None else None,
                                        starts.numel(), _stream())
        _lib.check("siIndexBuildDevice")
        assert rc == 0
        return self

    def build_host(self, starts, ends, values=None):
        s = np.ascontiguousarray(starts, np.int32)
        e = np.ascontiguousarray(ends, np.int32)
        v = None if values is None else np.ascontiguousarray(values, np.int32)
        if s.shape != e.shape:
            raise ValueError("starts and ends must have the same length")
        self._L.siIndexBuildHost(self._ix, s.ctypes.data, e.ctypes.data, None if v is None else v.ctypes.data, s.size)
        _lib.check("siIndexBuildHost")
        return self

    def export(self):
        """(starts, ends, values, branch[uint64, SI_NONE], perm) as numpy arrays."""
        n = len(self)
        s, e, v = (np.zeros(n, np.int32) for _ in range(3))
        b = np.zeros(n, np.uint64)
        p = np.zeros(n, np.uint32)
        if n:
            self._L.siIndexExport(self._ix, s.ctypes.data, e.ctypes.data, v.ctypes.data, b.ctypes.data, p.ctypes.data)
            _lib.check("siIndexExport")
        return s, e, v, b, p

    def view(self):
        dv = _lib.siDeviceView()
        self._L.siIndexDeviceView(self._ix, C.byref(dv))
        return dv

    # ---- queries ------------------------------------------------------------------------------
    def count(self, qs, qe, out=None, order=ORDER_AUTO):
        """Overlap counts (int32 tensor holding uint32 values) for device query tensors."""
        _chk_i32(qs, "qs"); _chk_i32(qe, "qe")
        if qs.numel() != qe.numel():
            raise ValueError("starts and ends must have the same length")
        n = qs.numel()
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=qs.device)
        self._L.siCountDevice(self._ix, qs.data_ptr(), qe.data_ptr(), n, out.data_ptr(), order, _stream())
        _lib.check("siCountDevice")
        return out

    def count_fanout(self, qs, qe, out, peer_ptrs, order=ORDER_AUTO):
        """count() that also stores every count at the same index of the arrays at `peer_ptrs` (raw device addresses,
        at most 15: other GPUs' copies of a gathered count vector -- siCountFanoutDevice, the fused count + all-gather)."""
        _chk_i32(qs, "qs"); _chk_i32(qe, "qe")
        n = qs.numel()
        arr = (C.c_void_p * max(1, len(peer_ptrs)))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        self._L.siCountFanoutDevice(self._ix, qs.data_ptr(), qe.data_ptr(), n, out.data_ptr(), arr, len(peer_ptrs), order, _stream())
        _lib.check("siCountFanoutDevice")
        return out

    def sort_queries(self, qs, qe):
        """Explicit first half of an ORDER_UNSORTED count (see siSortQueriesDevice)."""
        _chk_i32(qs, "qs"); _chk_i32(qe, "qe")
        self._L.siSortQueriesDevice(self._ix, qs.data_ptr(), qe.data_ptr(), qs.numel(), _stream())
        _lib.check("siSortQueriesDevice")

    def set_option(self, option, value):
        """siIndexSetOption: OPT_COUNT_ALGO (COUNT_AUTO/WALK/RANK/CELLS), OPT_BUCKET_INTERVALS, OPT_WINDOW_SHIFT."""
        rc = self._L.siIndexSetOption(self._ix, int(option), int(value))
        _lib.check("siIndexSetOption")
        if rc:
            raise ValueError(f"siIndexSetOption({option}, {value}) failed")
        return self

    def last_sort(self):
        """0: the input was already sorted; 1: narrow sort + tie fix; 2: composite 64-bit key (siIndexLastSort)."""
        return int(self._L.siIndexLastSort(self._ix))

    def cells_info(self):
        """Rank cells of the built index: {"starts": {...}, "ends": {...}, "pair": {...}} (siIndexCellsInfo)."""
        out = {}
        for which, name in ((0, "starts"), (1, "ends"), (2, "pair")):
            ci = _lib.siCellsInfo()
            if self._L.siIndexCellsInfo(self._ix, which, C.byref(ci)):
                raise RuntimeError("siIndexCellsInfo: index not built")
            out[name] = {"format": int(ci.format), "shift": int(ci.shift), "cells": int(ci.cells), "bytes": int(ci.bytes),
                         "overfull": int(ci.overfull), "direct": bool(ci.direct)}
        return out

    def bits_info(self):
        """Rank bits of the streaming count (siIndexBitsInfo): built, words, bytes, slow_words."""
        bi = _lib.siBitsInfo()
        if self._L.siIndexBitsInfo(self._ix, C.byref(bi)):
            raise RuntimeError("siIndexBitsInfo: index not built")
        return {"built": bool(bi.built), "words": int(bi.words), "bytes": int(bi.bytes), "slow_words": int(bi.slow_words)}

    def stream_stats(self):
        """(tiles, handed_back) of the last streaming count (siIndexStreamStats)."""
        t, f = C.c_ulonglong(0), C.c_ulonglong(0)
        self._L.siIndexStreamStats(self._ix, C.byref(t), C.byref(f))
        _lib.check("siIndexStreamStats")
        return int(t.value), int(f.value)

    def stab_info(self):
        """Stab lists of the CSR fill (siIndexStabInfo): state 0 not made yet / 1 in use / 2 the fill walks."""
        si = _lib.siStabInfo()
        if self._L.siIndexStabInfo(self._ix, C.byref(si)):
            raise RuntimeError("siIndexStabInfo: index not built")
        return {"state": int(si.state), "shift": int(si.shift), "lists": int(si.lists), "entries": int(si.entries),
                "record_bytes": int(si.record_bytes),
                "bytes": int(si.entries) * int(si.record_bytes) + (int(si.lists) + 1) * 8 if si.state == 1 else 0}

    def read_timings(self, max_records=4096):
        """[(kernel name, ms), ...] recorded since the last read (needs set_option(OPT_TIMING, 1))."""
        tags = (C.c_int * max_records)()
        ms = (C.c_float * max_records)()
        k = self._L.siIndexReadTimings(self._ix, tags, ms, max_records)
        return [(_lib.TAG_NAMES.get(tags[i], str(tags[i])), float(ms[i])) for i in range(k)]

    def has_overlaps(self, qs, qe):
        _chk_i32(qs, "qs"); _chk_i32(qe, "qe")
        out = torch.empty(qs.numel(), dtype=torch.uint8, device=qs.device)
        self._L.siAnyDevice(self._ix, qs.data_ptr(), qe.data_ptr(), qs.numel(), out.data_ptr(), _stream())
        _lib.check("siAnyDevice")
        return out.bool()

    def scan(self, counts, out=None):
        """Exclusive scan of counts -> int64 offsets[n+1] (offsets[n] = total hits)."""
        n = counts.numel()
        if out is None:
            out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
        self._L.siScanDevice(self._ix, counts.data_ptr(), n, out.data_ptr(), _stream())
        _lib.check("siScanDevice")
        return out

    def search(self, qs, qe, what=FILL_VALUES, order=ORDER_AUTO, counts=None, offsets=None, out=None):
        """count -> scan -> fill. Returns (offsets int64[n+1], flat results).

        what: FILL_VALUES int32[total], FILL_IDXS int32[total], FILL_KEYS int32[total,2],
        FILL_ITEMS int32[total,3]. One host sync to learn `total` unless `out` is given."""
        _chk_i32(qs, "qs"); _chk_i32(qe, "qe")
        n = qs.numel()
        if order == ORDER_AUTO and n:
            order = ORDER_SORTED if bool((qs[1:] >= qs[:-1]).all()) else ORDER_UNSORTED
        counts = self.count(qs, qe, out=counts, order=order)
        offsets = self.scan(counts, out=offsets)
        width = {FILL_VALUES: 1, FILL_IDXS: 1, FILL_KEYS: 2, FILL_ITEMS: 3}[what]
        if out is None:
            total = int(offsets[n].item())
            out = torch.empty((total, width) if width > 1 else (total,), dtype=torch.int32, device=qs.device)
        if n:
            self._L.siFillDevice(self._ix, qs.data_ptr(), qe.data_ptr(), n, offsets.data_ptr(), what,
                                 out.data_ptr(), order, _stream())
            _lib.check("siFillDevice")
        return offsets, out

    def search_values(self, qs, qe, order=ORDER_AUTO, **kw):
        return self.search(qs, qe, FILL_VALUES, order, **kw)

    def search_idxs(self, qs, qe, order=ORDER_AUTO, **kw):
        return self.search(qs, qe, FILL_IDXS, order, **kw)

    def search_keys(self, qs, qe, order=ORDER_AUTO, **kw):
        return self.search(qs, qe, FILL_KEYS, order, **kw)

    def coverage(self, qs, qe):
        _chk_i32(qs, "qs"); _chk_i32(qe, "qe")
        n = qs.numel()
        cnt = torch.empty(n, dtype=torch.int32, device=qs.device)
        cov = torch.empty(n, dtype=torch.int32, device=qs.device)
        self._L.siCoverageDevice(self._ix, qs.data_ptr(), qe.data_ptr(), n, cnt.data_ptr(), cov.data_ptr(), _stream())
        _lib.check("siCoverageDevice")
        return cnt, cov
