"""ctypes front-ends for the CPU checkers. TEST / BASELINE INFRASTRUCTURE ONLY.

* ``Oracle``     -> oracle/libsi_oracle.so  (plain-C restatement, si_oracle.c)
* ``Reference``  -> oracle/_ref/libsi_ref_{native,v3}.so (the UNMODIFIED reference
  C++ header behind oracle/ref_shim.cpp; compiled where /root/reference exists
  and shipped prebuilt to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module. The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NONE64 = np.uint64(0xFFFFFFFFFFFFFFFF)

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build_oracle(force: bool = False) -> str:
    """Compile oracle/si_oracle.c (gcc) if the .so is missing or stale."""
    so = os.path.join(HERE, "libsi_oracle.so")
    src = os.path.join(HERE, "si_oracle.c")
    src2 = os.path.join(HERE, "si_oracle_setops.c")
    hdr = os.path.join(HERE, "si_oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(p) > os.path.getmtime(so) for p in (src, src2, hdr))
    if force or stale:
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", src, src2, "-o", so])
    return so


class _IndexStruct(C.Structure):
    _fields_ = [("starts", C.POINTER(C.c_int32)), ("ends", C.POINTER(C.c_int32)),
                ("data", C.POINTER(C.c_int32)), ("branch", C.POINTER(C.c_size_t)),
                ("n", C.c_size_t), ("start_sorted_in", C.c_int), ("end_sorted_in", C.c_int)]


class Oracle:
    """Index + queries through the plain-C restatement (oracle/si_oracle.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(build_oracle())
            L.si_oracle_build.restype = C.POINTER(_IndexStruct)
            L.si_oracle_build.argtypes = [_i32p, _i32p, C.c_void_p, C.c_size_t]
            L.si_oracle_free.argtypes = [C.POINTER(_IndexStruct)]
            L.si_oracle_upper_bound.restype = C.c_size_t
            L.si_oracle_upper_bound.argtypes = [C.POINTER(_IndexStruct), C.c_int32]
            L.si_oracle_count_batch.argtypes = [C.POINTER(_IndexStruct), _i32p, _i32p, C.c_size_t, _u64p]
            L.si_oracle_has_overlaps_batch.argtypes = [C.POINTER(_IndexStruct), _i32p, _i32p, C.c_size_t, _u8p]
            L.si_oracle_search_batch.argtypes = [C.POINTER(_IndexStruct), _i32p, _i32p, C.c_size_t, _u64p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
            L.si_oracle_walk_stats.argtypes = [C.POINTER(_IndexStruct), _i32p, _i32p, C.c_size_t,
                                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            cls._lib = L
        return cls._lib

    def __init__(self, starts, ends, data=None):
        L = self.lib()
        s, e = _i32(starts), _i32(ends)
        assert s.shape == e.shape
        d = None if data is None else _i32(data)
        self._ix = L.si_oracle_build(s, e, None if d is None else d.ctypes.data, s.size)
        self.n = int(s.size)

    def __del__(self):
        if getattr(self, "_ix", None):
            self.lib().si_oracle_free(self._ix)
            self._ix = None

    # -- index arrays in position order -------------------------------------------------
    def _arr(self, field, dtype):
        if self.n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(getattr(self._ix.contents, field), shape=(self.n,)).astype(dtype)

    @property
    def starts(self): return self._arr("starts", np.int32)
    @property
    def ends(self): return self._arr("ends", np.int32)
    @property
    def data(self): return self._arr("data", np.int32)
    @property
    def branch(self): return self._arr("branch", np.uint64)

    # -- queries ------------------------------------------------------------------------
    def upper_bound(self, v):
        return int(self.lib().si_oracle_upper_bound(self._ix, int(v)))

    def count_batch(self, qs, qe):
        qs, qe = _i32(qs), _i32(qe)
        out = np.zeros(qs.size, np.uint64)
        self.lib().si_oracle_count_batch(self._ix, qs, qe, qs.size, out)
        return out

    def has_overlaps_batch(self, qs, qe):
        qs, qe = _i32(qs), _i32(qe)
        out = np.zeros(qs.size, np.uint8)
        self.lib().si_oracle_has_overlaps_batch(self._ix, qs, qe, qs.size, out)
        return out.astype(bool)

    def search_batch(self, qs, qe, want=("values",)):
        """Returns (offsets[nq+1] uint64, dict(values/idxs/keys))."""
        qs, qe = _i32(qs), _i32(qe)
        counts = self.count_batch(qs, qe)
        offsets = np.zeros(qs.size + 1, np.uint64)
        np.cumsum(counts, out=offsets[1:])
        total = int(offsets[-1])
        out = {}
        vals = np.zeros(total, np.int32) if "values" in want else None
        idxs = np.zeros(total, np.uint32) if "idxs" in want else None
        keys = np.zeros((total, 2), np.int32) if "keys" in want else None
        p = lambda a: None if a is None else a.ctypes.data
        self.lib().si_oracle_search_batch(self._ix, qs, qe, qs.size, offsets, p(vals), p(idxs), p(keys))
        if vals is not None: out["values"] = vals
        if idxs is not None: out["idxs"] = idxs
        if keys is not None: out["keys"] = keys
        return offsets, out

    def walk_stats(self, qs, qe):
        qs, qe = _i32(qs), _i32(qe)
        h, j = C.c_uint64(0), C.c_uint64(0)
        self.lib().si_oracle_walk_stats(self._ix, qs, qe, qs.size, C.byref(h), C.byref(j))
        return int(h.value), int(j.value)


# ---------------------------------------------------------------------------------------
def _host_cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


# gcc macro name -> /proc/cpuinfo flag, for the ISA extensions that matter to codegen
_MACRO2FLAG = {
    "AVX": "avx", "AVX2": "avx2", "FMA": "fma", "BMI": "bmi1", "BMI2": "bmi2", "POPCNT": "popcnt",
    "LZCNT": "abm", "MOVBE": "movbe", "F16C": "f16c", "ADX": "adx", "SHA": "sha_ni", "AES": "aes",
    "PCLMUL": "pclmulqdq", "VAES": "vaes", "VPCLMULQDQ": "vpclmulqdq", "GFNI": "gfni",
    "AVX512F": "avx512f", "AVX512BW": "avx512bw", "AVX512CD": "avx512cd", "AVX512DQ": "avx512dq",
    "AVX512VL": "avx512vl", "AVX512IFMA": "avx512ifma", "AVX512VBMI": "avx512vbmi",
    "AVX512VBMI2": "avx512_vbmi2", "AVX512VNNI": "avx512_vnni", "AVX512BITALG": "avx512_bitalg",
    "AVX512VPOPCNTDQ": "avx512_vpopcntdq", "AVX512BF16": "avx512_bf16", "AVX512FP16": "avx512_fp16",
    "AVXVNNI": "avx_vnni", "AMX_TILE": "amx_tile", "AMX_INT8": "amx_int8", "AMX_BF16": "amx_bf16",
    "SSE4_1": "sse4_1", "SSE4_2": "sse4_2", "CLWB": "clwb", "CLFLUSHOPT": "clflushopt",
    "FSGSBASE": "fsgsbase", "RDRND": "rdrand", "RDSEED": "rdseed", "SERIALIZE": "serialize",
    "MOVDIRI": "movdiri", "MOVDIR64B": "movdir64b", "CLDEMOTE": "cldemote", "PKU": "pku",
    "TSXLDTRK": "tsxldtrk", "XSAVE": "xsave", "XSAVEC": "xsavec", "XSAVEOPT": "xsaveopt",
    "XSAVES": "xsaves",
}


def reference_lib_path():
    """Pick the -march=native build when this host's ISA covers the build host's, else x86-64-v3.
    Returns (path, kind) or (None, reason)."""
    d = os.path.join(HERE, "_ref")
    native, v3 = os.path.join(d, "libsi_ref_native.so"), os.path.join(d, "libsi_ref_v3.so")
    flags_file = os.path.join(d, "cpuflags.txt")
    if os.path.exists(native) and os.path.exists(flags_file):
        need = {_MACRO2FLAG[m] for m in open(flags_file).read().split() if m in _MACRO2FLAG}
        if need and need <= _host_cpu_flags():
            return native, "native"
    if os.path.exists(v3) and "avx2" in _host_cpu_flags():
        return v3, "x86-64-v3"
    return None, "oracle/_ref not built (reference absent at build time)"


class Reference:
    """The real reference (si::IntervalMap<int,int>) through oracle/ref_shim.cpp."""

    _lib = None
    kind = None

    @classmethod
    def available(cls):
        return reference_lib_path()[0] is not None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path, kind = reference_lib_path()
            if path is None:
                raise RuntimeError(kind)
            L = C.CDLL(path)
            cls.kind = kind
            vp = C.c_void_p
            L.si_ref_create.restype = vp
            L.si_ref_destroy.argtypes = [vp]
            L.si_ref_clear.argtypes = [vp]
            L.si_ref_add.argtypes = [vp, C.c_int, C.c_int, C.c_int]
            L.si_ref_add_many.argtypes = [vp, _i32p, _i32p, C.c_void_p, C.c_size_t]
            L.si_ref_build.argtypes = [vp]
            L.si_ref_size.restype = C.c_size_t
            L.si_ref_size.argtypes = [vp]
            L.si_ref_flags.argtypes = [vp]
            L.si_ref_export.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.si_ref_upper_bound.restype = C.c_uint64
            L.si_ref_upper_bound.argtypes = [vp, C.c_int]
            L.si_ref_count.restype = C.c_uint64
            L.si_ref_count.argtypes = [vp, C.c_int, C.c_int]
            L.si_ref_has_overlaps.argtypes = [vp, C.c_int, C.c_int]
            L.si_ref_count_batch.argtypes = [vp, _i32p, _i32p, C.c_size_t, _u64p, C.c_int, C.c_int]
            L.si_ref_has_overlaps_batch.argtypes = [vp, _i32p, _i32p, C.c_size_t, _u8p]
            for name in ("si_ref_search_values_batch",):
                f = getattr(L, name)
                f.restype = C.c_uint64
                f.argtypes = [vp, _i32p, _i32p, C.c_size_t, _u64p, C.c_int, C.c_int]
            for name in ("si_ref_search_keys_batch", "si_ref_search_idxs_batch"):
                f = getattr(L, name)
                f.restype = C.c_uint64
                f.argtypes = [vp, _i32p, _i32p, C.c_size_t, _u64p]
            L.si_ref_take_values.argtypes = [vp, C.c_void_p]
            L.si_ref_take_keys.argtypes = [vp, C.c_void_p]
            L.si_ref_take_idxs.argtypes = [vp, C.c_void_p]
            L.si_ref_coverage_batch.argtypes = [vp, _i32p, _i32p, C.c_size_t, _u64p, _i32p]
            for name in ("si_ref_time_count", "si_ref_time_search_values"):
                f = getattr(L, name)
                f.restype = C.c_double
                f.argtypes = [vp, _i32p, _i32p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]
            L.si_ref_time_build.restype = C.c_double
            L.si_ref_time_build.argtypes = [vp, _i32p, _i32p, C.c_size_t]
            cls._lib = L
        return cls._lib

    def __init__(self, starts=None, ends=None, data=None, build=True):
        self._h = self.lib().si_ref_create()
        if starts is not None:
            self.add_many(starts, ends, data)
            if build:
                self.build()

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().si_ref_destroy(self._h)
            self._h = None

    def add(self, s, e, v): self.lib().si_ref_add(self._h, int(s), int(e), int(v))

    def add_many(self, starts, ends, data=None):
        s, e = _i32(starts), _i32(ends)
        d = None if data is None else _i32(data)
        self.lib().si_ref_add_many(self._h, s, e, None if d is None else d.ctypes.data, s.size)

    def build(self): self.lib().si_ref_build(self._h)
    def size(self): return int(self.lib().si_ref_size(self._h))
    def flags(self): return int(self.lib().si_ref_flags(self._h))

    def export(self):
        n = self.size()
        s, e, d = (np.zeros(n, np.int32) for _ in range(3))
        b = np.full(n, NONE64, np.uint64)
        self.lib().si_ref_export(self._h, s.ctypes.data, e.ctypes.data, d.ctypes.data, b.ctypes.data)
        return s, e, d, b

    def upper_bound(self, v): return int(self.lib().si_ref_upper_bound(self._h, int(v)))
    def count(self, s, e): return int(self.lib().si_ref_count(self._h, int(s), int(e)))
    def has_overlaps(self, s, e): return bool(self.lib().si_ref_has_overlaps(self._h, int(s), int(e)))

    def count_batch(self, qs, qe, mode=0, threads=1):
        qs, qe = _i32(qs), _i32(qe)
        out = np.zeros(qs.size, np.uint64)
        self.lib().si_ref_count_batch(self._h, qs, qe, qs.size, out, mode, threads)
        return out

    def has_overlaps_batch(self, qs, qe):
        qs, qe = _i32(qs), _i32(qe)
        out = np.zeros(qs.size, np.uint8)
        self.lib().si_ref_has_overlaps_batch(self._h, qs, qe, qs.size, out)
        return out.astype(bool)

    def search_values_batch(self, qs, qe, mode=0, threads=1):
        qs, qe = _i32(qs), _i32(qe)
        off = np.zeros(qs.size + 1, np.uint64)
        total = self.lib().si_ref_search_values_batch(self._h, qs, qe, qs.size, off, mode, threads)
        vals = np.zeros(int(total), np.int32)
        self.lib().si_ref_take_values(self._h, vals.ctypes.data)
        return off, vals

    def search_keys_batch(self, qs, qe):
        qs, qe = _i32(qs), _i32(qe)
        off = np.zeros(qs.size + 1, np.uint64)
        total = self.lib().si_ref_search_keys_batch(self._h, qs, qe, qs.size, off)
        keys = np.zeros((int(total), 2), np.int32)
        self.lib().si_ref_take_keys(self._h, keys.ctypes.data)
        return off, keys

    def search_idxs_batch(self, qs, qe):
        qs, qe = _i32(qs), _i32(qe)
        off = np.zeros(qs.size + 1, np.uint64)
        total = self.lib().si_ref_search_idxs_batch(self._h, qs, qe, qs.size, off)
        idx = np.zeros(int(total), np.uint64)
        self.lib().si_ref_take_idxs(self._h, idx.ctypes.data)
        return off, idx

    def coverage_batch(self, qs, qe):
        qs, qe = _i32(qs), _i32(qe)
        c = np.zeros(qs.size, np.uint64)
        v = np.zeros(qs.size, np.int32)
        self.lib().si_ref_coverage_batch(self._h, qs, qe, qs.size, c, v)
        return c, v

    # timing legs (seconds, found-checksum)
    def time_count(self, qs, qe, threads=1):
        qs, qe = _i32(qs), _i32(qe)
        f = C.c_uint64(0)
        t = self.lib().si_ref_time_count(self._h, qs, qe, qs.size, threads, C.byref(f))
        return float(t), int(f.value)

    def time_search_values(self, qs, qe, threads=1):
        qs, qe = _i32(qs), _i32(qe)
        f = C.c_uint64(0)
        t = self.lib().si_ref_time_search_values(self._h, qs, qe, qs.size, threads, C.byref(f))
        return float(t), int(f.value)

    def time_build(self, starts, ends):
        s, e = _i32(starts), _i32(ends)
        return float(self.lib().si_ref_time_build(self._h, s, e, s.size))


# ---------------------------------------------------------------------------------------
# set algebra (oracle/si_oracle_setops.c) and the compiled reference C header beside it
# ---------------------------------------------------------------------------------------
class _SetStruct(C.Structure):
    _fields_ = [("starts", C.POINTER(C.c_int32)), ("ends", C.POINTER(C.c_int32)), ("data", C.POINTER(C.c_int32)),
                ("n", C.c_size_t), ("cap", C.c_size_t)]


COMBINE_T = C.CFUNCTYPE(C.c_int32, C.c_int32, C.c_int32)
COMBINERS = {None: None,
             "sum": COMBINE_T(lambda a, b: C.c_int32(a + b).value),
             "max": COMBINE_T(lambda a, b: max(a, b)),
             "second": COMBINE_T(lambda a, b: b)}
SETOPS = ("merge", "gaps", "union", "intersection", "difference", "symmetric_difference", "span", "expand", "flank", "unique")


def _triple(s, e, d=None):
    s, e = _i32(s), _i32(e)
    d = np.arange(s.size, dtype=np.int32) if d is None else _i32(d)
    return s, e, d


class OracleSetOps:
    """The set-algebra restatement over plain arrays. Every method returns (starts, ends, data) int32
    arrays in the reference's emission order (span: (lo, hi) or None)."""

    _bound = False

    @classmethod
    def lib(cls):
        L = Oracle.lib()
        if not cls._bound:
            P, ip, sz, i32, vp = C.POINTER(_SetStruct), _i32p, C.c_size_t, C.c_int32, C.c_void_p
            IX = C.POINTER(_IndexStruct)
            L.si_oracle_set_free.argtypes = [P]
            for name, args in (("merge", [ip, ip, ip, sz, vp]), ("gaps", [ip, ip, ip, sz, i32, i32, i32]),
                               ("union", [ip, ip, ip, sz, ip, ip, ip, sz, vp]), ("intersection", [ip, ip, ip, sz, IX, vp]),
                               ("difference", [ip, ip, ip, sz, IX]), ("symmetric_difference", [IX, IX]),
                               ("expand", [ip, ip, ip, sz, i32, i32, i32, i32]), ("flank", [ip, ip, ip, sz, i32, i32, i32, i32]),
                               ("unique", [ip, ip, ip, sz, vp])):
                f = getattr(L, "si_oracle_" + name)
                f.restype, f.argtypes = P, args
            L.si_oracle_span.restype = C.c_int
            L.si_oracle_span.argtypes = [ip, ip, sz, C.POINTER(i32), C.POINTER(i32)]
            cls._bound = True
        return L

    @classmethod
    def _take(cls, p):
        n = int(p.contents.n)
        out = tuple(np.ctypeslib.as_array(getattr(p.contents, f), shape=(n,)).copy() if n else np.zeros(0, np.int32)
                    for f in ("starts", "ends", "data"))
        cls.lib().si_oracle_set_free(p)
        return out

    @staticmethod
    def _fn(combine):
        f = COMBINERS[combine]
        return None if f is None else C.cast(f, C.c_void_p)

    @classmethod
    def merge(cls, s, e, d=None, combine=None):
        s, e, d = _triple(s, e, d)
        return cls._take(cls.lib().si_oracle_merge(s, e, d, s.size, cls._fn(combine)))

    @classmethod
    def gaps(cls, s, e, d, lo, hi, fill):
        s, e, d = _triple(s, e, d)
        return cls._take(cls.lib().si_oracle_gaps(s, e, d, s.size, lo, hi, fill))

    @classmethod
    def union(cls, a, b, combine=None):
        a, b = _triple(*a), _triple(*b)
        return cls._take(cls.lib().si_oracle_union(a[0], a[1], a[2], a[0].size, b[0], b[1], b[2], b[0].size, cls._fn(combine)))

    @classmethod
    def intersection(cls, a, b, combine=None):
        """a: stored order arrays; b: arrays of the OTHER set, indexed here (payload = its data)."""
        a, b = _triple(*a), _triple(*b)
        o = Oracle(b[0], b[1], b[2])
        return cls._take(cls.lib().si_oracle_intersection(a[0], a[1], a[2], a[0].size, o._ix, cls._fn(combine)))

    @classmethod
    def difference(cls, a, b):
        a, b = _triple(*a), _triple(*b)
        o = Oracle(b[0], b[1], b[2])
        return cls._take(cls.lib().si_oracle_difference(a[0], a[1], a[2], a[0].size, o._ix))

    @classmethod
    def symmetric_difference(cls, a, b):
        a, b = _triple(*a), _triple(*b)
        oa, ob = Oracle(a[0], a[1], a[2]), Oracle(b[0], b[1], b[2])
        return cls._take(cls.lib().si_oracle_symmetric_difference(oa._ix, ob._ix))

    @classmethod
    def span(cls, s, e):
        s, e = _i32(s), _i32(e)
        lo, hi = C.c_int32(0), C.c_int32(0)
        return (lo.value, hi.value) if cls.lib().si_oracle_span(s, e, s.size, C.byref(lo), C.byref(hi)) else None

    @classmethod
    def expand(cls, s, e, d, left, right, lo, hi):
        s, e, d = _triple(s, e, d)
        return cls._take(cls.lib().si_oracle_expand(s, e, d, s.size, left, right, lo, hi))

    @classmethod
    def flank(cls, s, e, d, left, right, lo, hi):
        s, e, d = _triple(s, e, d)
        return cls._take(cls.lib().si_oracle_flank(s, e, d, s.size, left, right, lo, hi))

    @classmethod
    def unique(cls, s, e, d=None, combine=None):
        s, e, d = _triple(s, e, d)
        return cls._take(cls.lib().si_oracle_unique(s, e, d, s.size, cls._fn(combine)))


class CSetOps:
    """The same operations through a library exporting the reference's C ABI: the compiled reference
    header (oracle/_ref/libsi_cref.so, `CSetOps.reference()`) or libsuperintervals_b200.so
    (`CSetOps(lib)`). Sets are created with addInterval x n in stored order; `other` sets are indexed."""

    CREF = os.path.join(HERE, "_ref", "libsi_cref.so")

    class _SI(C.Structure):
        _fields_ = [("starts", C.POINTER(C.c_int32)), ("ends", C.POINTER(C.c_int32)), ("data", C.POINTER(C.c_int32)),
                    ("branch", C.POINTER(C.c_size_t)), ("size", C.c_size_t), ("capacity", C.c_size_t), ("idx", C.c_size_t),
                    ("startSorted", C.c_bool), ("endSorted", C.c_bool)]

    @classmethod
    def reference_available(cls):
        return os.path.exists(cls.CREF)

    @classmethod
    def reference(cls):
        return cls(C.CDLL(cls.CREF))

    def __init__(self, L):
        # a library already bound elsewhere (superintervals_b200._lib.bind) keeps its handle type
        rt = L.createSuperIntervals.restype
        SI = rt if isinstance(rt, type) and issubclass(rt, C._Pointer) else C.POINTER(self._SI)
        self.SI = SI
        i32, vp = C.c_int32, C.c_void_p
        self.L = L
        L.createSuperIntervals.restype = SI
        L.destroySuperIntervals.argtypes = [SI]
        L.addInterval.argtypes = [SI, i32, i32, i32]
        L.indexSuperIntervals.argtypes = [SI]
        for name, args in (("mergeOverlaps", [SI, vp]), ("intervalGaps", [SI, i32, i32, i32]), ("unionWith", [SI, SI, vp]),
                           ("intersection", [SI, SI, vp]), ("difference", [SI, SI]), ("symmetricDifference", [SI, SI]),
                           ("expandIntervals", [SI, i32, i32, i32, i32]), ("flankIntervals", [SI, i32, i32, i32, i32]),
                           ("uniqueIntervals", [SI, vp])):
            f = getattr(L, name)
            f.restype, f.argtypes = SI, args
        L.intervalSpan.restype = C.c_bool
        L.intervalSpan.argtypes = [SI, C.POINTER(i32), C.POINTER(i32)]

    def make(self, s, e, d=None, index=False):
        s, e, d = _triple(s, e, d)
        si = self.L.createSuperIntervals()
        if hasattr(self.L, "addIntervals"):
            self.L.addIntervals.argtypes = [self.SI, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
            self.L.addIntervals(si, s.ctypes.data, e.ctypes.data, d.ctypes.data, s.size)
        else:
            for k in range(s.size):
                self.L.addInterval(si, int(s[k]), int(e[k]), int(d[k]))
        if index:
            self.L.indexSuperIntervals(si)
        return si

    def take(self, si, flags=False):
        n = int(si.contents.size)
        out = tuple(np.ctypeslib.as_array(getattr(si.contents, f), shape=(n,)).copy() if n else np.zeros(0, np.int32)
                    for f in ("starts", "ends", "data"))
        if flags:
            out = out + ((bool(si.contents.startSorted), bool(si.contents.endSorted)),)
        self.L.destroySuperIntervals(si)
        return out

    @staticmethod
    def _fn(combine):
        f = COMBINERS[combine]
        return None if f is None else C.cast(f, C.c_void_p)

    def run(self, op, a, b=None, combine=None, args=(), index_a=False, flags=False):
        """op in SETOPS; a, b = (starts, ends[, data]) in stored order; returns arrays (span: (lo, hi) or None)."""
        A = self.make(*a, index=index_a or op == "symmetric_difference")
        B = self.make(*b, index=op in ("intersection", "difference", "symmetric_difference")) if b is not None else None
        L = self.L
        try:
            if op == "span":
                lo, hi = C.c_int32(0), C.c_int32(0)
                return (lo.value, hi.value) if L.intervalSpan(A, C.byref(lo), C.byref(hi)) else None
            r = {"merge": lambda: L.mergeOverlaps(A, self._fn(combine)), "gaps": lambda: L.intervalGaps(A, *args),
                 "union": lambda: L.unionWith(A, B, self._fn(combine)), "intersection": lambda: L.intersection(A, B, self._fn(combine)),
                 "difference": lambda: L.difference(A, B), "symmetric_difference": lambda: L.symmetricDifference(A, B),
                 "expand": lambda: L.expandIntervals(A, *args), "flank": lambda: L.flankIntervals(A, *args),
                 "unique": lambda: L.uniqueIntervals(A, self._fn(combine))}[op]()
            return self.take(r, flags)
        finally:
            L.destroySuperIntervals(A)
            if B is not None:
                L.destroySuperIntervals(B)
