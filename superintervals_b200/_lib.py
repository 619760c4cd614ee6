"""ctypes binding of libsuperintervals_b200.so (the C ABI in include/*.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C superintervals_b200/csrc``. There is no fallback: if it is missing,
importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsuperintervals_b200.so")
SI_NONE = (1 << 64) - 1

ORDER_AUTO, ORDER_SORTED, ORDER_UNSORTED, ORDER_ASIS = 0, 1, 2, 3
FILL_VALUES, FILL_IDXS, FILL_KEYS, FILL_ITEMS = 0, 1, 2, 3
OPT_COUNT_ALGO, OPT_BUCKET_INTERVALS, OPT_WINDOW_SHIFT, OPT_TIMING, OPT_GRID_INTERVALS = 0, 1, 2, 3, 4
OPT_CELLS_DIRECT_BYTES, OPT_CELLS_FILL = 5, 6
OPT_STAB_LISTS, OPT_STAB_BUDGET = 7, 8
OPT_STREAM, OPT_STREAM_BUDGET, OPT_L2_PERSIST, OPT_NARROW_SORT, OPT_RESIDENT_QUERIES, OPT_STAB_VALUE_LISTS = 9, 10, 11, 12, 13, 14
OPT_PAIR_CELLS = 15
TAG_NAMES = {1: "pt_histogram", 2: "pt_onesweep", 3: "count_walk", 4: "count_rank", 5: "scan", 6: "fill", 7: "count_cells", 8: "fill_runs", 9: "count_stream"}
COUNT_AUTO, COUNT_WALK, COUNT_RANK, COUNT_CELLS = 0, 1, 2, 3


class cSuperIntervals(C.Structure):
    """include/c_superintervals.h (reference c_superintervals.h:81-91)."""
    _fields_ = [("starts", C.POINTER(C.c_int32)), ("ends", C.POINTER(C.c_int32)),
                ("data", C.POINTER(C.c_int32)), ("branch", C.POINTER(C.c_size_t)),
                ("size", C.c_size_t), ("capacity", C.c_size_t), ("idx", C.c_size_t),
                ("startSorted", C.c_bool), ("endSorted", C.c_bool)]


class cIndexResult(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_int32)), ("size", C.c_size_t), ("capacity", C.c_size_t)]


class KeyPair(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32)]


class Interval(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32), ("data", C.c_int32)]


class cKeyResult(C.Structure):
    _fields_ = [("data", C.POINTER(KeyPair)), ("size", C.c_size_t), ("capacity", C.c_size_t)]


class cItemResult(C.Structure):
    _fields_ = [("data", C.POINTER(Interval)), ("size", C.c_size_t), ("capacity", C.c_size_t)]


class siDeviceView(C.Structure):
    _fields_ = [("starts", C.c_void_p), ("ends", C.c_void_p), ("values", C.c_void_p),
                ("branch", C.c_void_p), ("n", C.c_size_t), ("device", C.c_int)]


class siCellsInfo(C.Structure):
    _fields_ = [("format", C.c_uint), ("shift", C.c_uint), ("cells", C.c_ulonglong), ("bytes", C.c_ulonglong),
                ("overfull", C.c_ulonglong), ("direct", C.c_int)]


class siBitsInfo(C.Structure):
    _fields_ = [("built", C.c_int), ("words", C.c_ulonglong), ("bytes", C.c_ulonglong), ("slow_words", C.c_ulonglong)]


class siMultiStats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_count", C.c_double), ("ms_gather", C.c_double),
                ("ms_d2h", C.c_double), ("nccl_bytes", C.c_ulonglong), ("nccl_version", C.c_int),
                ("peer_access", C.c_int), ("peer_bytes", C.c_ulonglong)]


class siBedTable(C.Structure):
    _fields_ = [("contig", C.POINTER(C.c_int32)), ("starts", C.POINTER(C.c_int32)), ("ends", C.POINTER(C.c_int32)),
                ("n", C.c_size_t), ("lines", C.c_size_t), ("skipped", C.c_size_t),
                ("names", C.POINTER(C.c_char_p)), ("n_contigs", C.c_size_t), ("contig_offsets", C.POINTER(C.c_size_t))]


class siStabInfo(C.Structure):
    _fields_ = [("state", C.c_int), ("shift", C.c_uint), ("lists", C.c_ulonglong), ("entries", C.c_ulonglong),
                ("record_bytes", C.c_uint)]


def build_library(verbose: bool = False) -> str:
    """nvcc-compile the CUDA library for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout + out.stderr)
    if out.returncode:
        raise RuntimeError("building libsuperintervals_b200.so failed")
    return LIB_PATH


# every symbol include/c_superintervals.h and include/superintervals_b200.h declare
C_ABI_SYMBOLS = [
    "createSuperIntervals", "destroySuperIntervals", "clearSuperIntervals", "reserveSuperIntervals",
    "addInterval", "sizeSuperIntervals", "sortIntervals", "indexSuperIntervals", "intervalAt", "startAt",
    "endAt", "dataAt", "upperBound", "anyOverlaps", "countOverlaps", "searchValues", "searchIdxs",
    "searchKeys", "searchItems", "searchPoint", "coverage", "findOverlaps", "createIndexResult",
    "clearIndexResult", "destroyIndexResult", "createKeyResult", "clearKeyResult", "destroyKeyResult",
    "createItemResult", "clearItemResult", "destroyItemResult",
    "mergeOverlaps", "intervalGaps", "unionWith", "intersection", "difference", "symmetricDifference",
    "intervalSpan", "expandIntervals", "flankIntervals", "uniqueIntervals",
]
B200_SYMBOLS = [
    "si_b200_last_error", "si_b200_last_error_string", "si_b200_clear_error", "si_b200_version",
    "si_b200_device_count", "si_b200_kernel_launches", "addIntervals", "siSetHostMirror",
    "countOverlapsBatch", "countOverlapsBatch32", "anyOverlapsBatch", "searchValuesBatch", "searchIdxsBatch", "searchKeysBatch",
    "searchItemsBatch", "coverageBatch", "intersectionPairs", "siParseBed", "siBedTableFree", "siIndexCreate", "siIndexDestroy", "siIndexOf", "siIndexSize",
    "siIndexDeviceView", "siIndexBuildHost", "siIndexBuildDevice", "siIndexExport", "siCountDevice",
    "siCountDevice64", "siSortQueriesDevice", "siIndexSetOption", "siIndexCellsInfo", "siIndexBitsInfo", "siIndexStreamStats", "siIndexStabInfo", "siIndexReadTimings", "siAnyDevice", "siScanDevice", "siFillDevice", "siCoverageDevice",
    "siIndexDeviceBytes", "siRouteByContigDevice", "siScatterCountsDevice", "siCountMixedDevice", "siCountMixedPeerDevice", "siIndexLastSort", "siCountFanoutDevice", "siPeerBarrierDevice", "siIpcAlloc", "siIpcOpen", "siIpcClose", "siIpcFree",
    "siMultiCreate", "siMultiDestroy", "siMultiDeviceCount", "siMultiIndexOf", "siMultiBuildReplicated", "siMultiCountBatch",
    "siMultiSearchValuesBatch", "siMultiDeviceCounts", "siMultiLastStats",
]

_lib = None


def bind(L):
    """Attach argtypes/restypes for the C ABI to a CDLL. Works for any library exporting the
    reference's C ABI (tests point it at the reference build in oracle/_ref as well)."""
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    SI = C.POINTER(cSuperIntervals)
    L.createSuperIntervals.restype = SI
    L.createSuperIntervals.argtypes = []
    L.destroySuperIntervals.argtypes = [SI]
    L.clearSuperIntervals.argtypes = [SI]
    L.reserveSuperIntervals.argtypes = [SI, sz]
    L.addInterval.argtypes = [SI, i32, i32, i32]
    L.sizeSuperIntervals.restype = sz
    L.sizeSuperIntervals.argtypes = [SI]
    L.sortIntervals.argtypes = [SI]
    L.indexSuperIntervals.argtypes = [SI]
    L.intervalAt.restype = C.c_bool
    L.intervalAt.argtypes = [SI, sz, C.POINTER(Interval)]
    for f in (L.startAt, L.endAt, L.dataAt):
        f.restype = i32
        f.argtypes = [SI, sz]
    L.upperBound.restype = sz
    L.upperBound.argtypes = [SI, i32]
    L.anyOverlaps.restype = C.c_bool
    L.anyOverlaps.argtypes = [SI, i32, i32]
    L.countOverlaps.restype = sz
    L.countOverlaps.argtypes = [SI, i32, i32]
    L.searchValues.argtypes = [SI, i32, i32, C.POINTER(cIndexResult)]
    L.searchIdxs.argtypes = [SI, i32, i32, C.POINTER(cIndexResult)]
    L.searchKeys.argtypes = [SI, i32, i32, C.POINTER(cKeyResult)]
    L.searchItems.argtypes = [SI, i32, i32, C.POINTER(cItemResult)]
    L.searchPoint.argtypes = [SI, i32, C.POINTER(cIndexResult)]
    L.coverage.argtypes = [SI, i32, i32, C.POINTER(sz), C.POINTER(i32)]
    L.findOverlaps.argtypes = [SI, i32, i32, C.POINTER(i32), C.POINTER(sz)]
    # set operations (c_superintervals.h:243-334): new un-indexed handles
    for name, args in (("mergeOverlaps", [SI, vp]), ("intervalGaps", [SI, i32, i32, i32]), ("unionWith", [SI, SI, vp]),
                       ("intersection", [SI, SI, vp]), ("difference", [SI, SI]), ("symmetricDifference", [SI, SI]),
                       ("expandIntervals", [SI, i32, i32, i32, i32]), ("flankIntervals", [SI, i32, i32, i32, i32]),
                       ("uniqueIntervals", [SI, vp])):
        if hasattr(L, name):
            getattr(L, name).restype = SI
            getattr(L, name).argtypes = args
    if hasattr(L, "intervalSpan"):
        L.intervalSpan.restype = C.c_bool
        L.intervalSpan.argtypes = [SI, C.POINTER(i32), C.POINTER(i32)]
    L.createIndexResult.restype = cIndexResult
    L.createKeyResult.restype = cKeyResult
    L.createItemResult.restype = cItemResult
    for f, T in ((L.clearIndexResult, cIndexResult), (L.destroyIndexResult, cIndexResult),
                 (L.clearKeyResult, cKeyResult), (L.destroyKeyResult, cKeyResult),
                 (L.clearItemResult, cItemResult), (L.destroyItemResult, cItemResult)):
        f.argtypes = [C.POINTER(T)]
    return L


def bind_b200(L):
    vp, sz = C.c_void_p, C.c_size_t
    SI = C.POINTER(cSuperIntervals)
    L.si_b200_last_error.restype = C.c_int
    L.si_b200_last_error_string.restype = C.c_char_p
    L.si_b200_version.restype = C.c_char_p
    L.si_b200_device_count.restype = C.c_int
    L.si_b200_kernel_launches.restype = C.c_ulonglong
    L.addIntervals.argtypes = [SI, vp, vp, vp, sz]
    L.siSetHostMirror.argtypes = [SI, C.c_bool]
    L.countOverlapsBatch.argtypes = [SI, vp, vp, sz, vp]
    L.countOverlapsBatch32.argtypes = [SI, vp, vp, sz, vp]
    L.anyOverlapsBatch.argtypes = [SI, vp, vp, sz, vp]
    L.searchValuesBatch.argtypes = [SI, vp, vp, sz, vp, C.POINTER(cIndexResult)]
    L.searchIdxsBatch.argtypes = [SI, vp, vp, sz, vp, C.POINTER(cIndexResult)]
    L.searchKeysBatch.argtypes = [SI, vp, vp, sz, vp, C.POINTER(cKeyResult)]
    L.searchItemsBatch.argtypes = [SI, vp, vp, sz, vp, C.POINTER(cItemResult)]
    L.coverageBatch.argtypes = [SI, vp, vp, sz, vp, vp]
    L.siIndexCreate.restype = vp
    L.siIndexDestroy.argtypes = [vp]
    L.siIndexOf.restype = vp
    L.siIndexOf.argtypes = [SI]
    L.siIndexSize.restype = sz
    L.siIndexSize.argtypes = [vp]
    L.siIndexDeviceBytes.restype = sz
    L.siIndexDeviceBytes.argtypes = [vp]
    L.siIndexDeviceView.argtypes = [vp, C.POINTER(siDeviceView)]
    L.siIndexBuildHost.argtypes = [vp, vp, vp, vp, sz]
    L.siIndexBuildDevice.argtypes = [vp, vp, vp, vp, sz, vp]
    L.siIndexExport.argtypes = [vp, vp, vp, vp, vp, vp]
    L.siCountDevice.argtypes = [vp, vp, vp, sz, vp, C.c_int, vp]
    L.siCountDevice64.argtypes = [vp, vp, vp, sz, vp, C.c_int, vp]
    L.siSortQueriesDevice.argtypes = [vp, vp, vp, sz, vp]
    L.siIndexCellsInfo.argtypes = [vp, C.c_int, C.POINTER(siCellsInfo)]
    L.siIndexCellsInfo.restype = C.c_int
    L.siMultiCreate.restype = vp
    L.siMultiCreate.argtypes = [vp, C.c_int]
    L.siMultiDestroy.argtypes = [vp]
    L.siMultiDeviceCount.argtypes = [vp]
    L.siMultiIndexOf.restype = vp
    L.siMultiIndexOf.argtypes = [vp, C.c_int]
    L.siMultiBuildReplicated.argtypes = [vp, vp, vp, vp, sz]
    L.siMultiCountBatch.argtypes = [vp, vp, vp, sz, vp]
    L.siMultiSearchValuesBatch.argtypes = [vp, vp, vp, sz, vp, C.POINTER(cIndexResult)]
    L.siMultiDeviceCounts.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    L.siMultiLastStats.argtypes = [vp, C.POINTER(siMultiStats)]
    L.siIndexBitsInfo.argtypes = [vp, C.POINTER(siBitsInfo)]
    L.siIndexBitsInfo.restype = C.c_int
    L.siIndexStreamStats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.siIndexStreamStats.restype = C.c_int
    if hasattr(L, "intersectionPairs"):
        L.intersectionPairs.restype = SI
        L.intersectionPairs.argtypes = [SI, SI, C.POINTER(cIndexResult)]
    L.siParseBed.restype = C.c_int
    L.siParseBed.argtypes = [C.c_char_p, sz, C.c_int, C.c_int, C.c_int, C.POINTER(siBedTable)]
    L.siBedTableFree.argtypes = [C.POINTER(siBedTable)]
    L.siIndexStabInfo.argtypes = [vp, C.POINTER(siStabInfo)]
    L.siIndexStabInfo.restype = C.c_int
    L.siIndexSetOption.argtypes = [vp, C.c_int, C.c_longlong]
    L.siIndexSetOption.restype = C.c_int
    L.siIndexReadTimings.argtypes = [vp, vp, vp, C.c_int]
    L.siIndexReadTimings.restype = C.c_int
    L.siAnyDevice.argtypes = [vp, vp, vp, sz, vp, vp]
    L.siScanDevice.argtypes = [vp, vp, sz, vp, vp]
    L.siFillDevice.argtypes = [vp, vp, vp, sz, vp, C.c_int, vp, C.c_int, vp]
    L.siCoverageDevice.argtypes = [vp, vp, vp, sz, vp, vp, vp]
    L.siRouteByContigDevice.argtypes = [vp, vp, vp, vp, sz, C.c_int, vp, vp, vp, vp, vp]
    L.siScatterCountsDevice.argtypes = [vp, vp, vp, sz, vp, vp]
    L.siIndexLastSort.argtypes = [vp]
    L.siCountFanoutDevice.argtypes = [vp, vp, vp, sz, vp, vp, C.c_int, C.c_int, vp]
    L.siPeerBarrierDevice.argtypes = [vp, vp, C.c_int, C.c_uint32, vp, vp]
    L.siIpcAlloc.argtypes = [sz, C.POINTER(vp), vp]
    L.siIpcOpen.argtypes = [vp, C.POINTER(vp)]
    L.siIpcClose.argtypes = [vp]
    L.siIpcFree.argtypes = [vp]
    L.siCountMixedDevice.argtypes = [vp, C.c_int, vp, vp, vp, sz, vp, vp, vp]
    L.siCountMixedPeerDevice.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    return L


def lib():
    """The loaded product library. Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C superintervals_b200/csrc`. superintervals_b200 has no CPU fallback.")
        # SIB_LIBRARY: an alternative build of the same sources (kernel tuning experiments, tools/variants.sh)
        _lib = bind_b200(bind(C.CDLL(os.environ.get("SIB_LIBRARY") or LIB_PATH)))
    return _lib


class SuperIntervalsError(RuntimeError):
    pass


def check(where: str = ""):
    """Raise if the library latched a CUDA error (the C ABI itself cannot report one)."""
    L = lib()
    code = L.si_b200_last_error()
    if code:
        msg = L.si_b200_last_error_string().decode()
        L.si_b200_clear_error()
        raise SuperIntervalsError(f"{where}: {msg}" if where else msg)
