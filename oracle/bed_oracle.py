"""CPU restatement of the reference's BED tokenising (TEST INFRASTRUCTURE, small inputs only).

Follows reference test/bench.cpp:67-102: lines by std::getline, fields by '\\t', numbers by
std::stoi (leading whitespace, optional sign, digits, trailing junk ignored), and
examples/bed-intersect-si.rs:63-123 for keeping every chrom (bench.cpp keeps "chr1" only -- a
filter on the chrom column of this output). Where std::stoi would throw (no digits / out of
range) or a field is missing, the line is skipped and counted, as csrc/bed.cu documents.
"""
import re

import numpy as np

_STOI = re.compile(rb"^[ \t\n\v\f\r]*([+-]?[0-9]+)")
I32 = np.iinfo(np.int32)


def _stoi(tok):
    m = _STOI.match(tok)
    if not m:
        return None
    v = int(m.group(1))
    return v if I32.min <= v <= I32.max else None


def parse_bed(text: bytes, normalize=False, end_shift=0):
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                      # std::getline yields nothing after a final newline
    names, ids, contig, starts, ends = [], {}, [], [], []
    for ln in lines:
        f = ln.split(b"\t")
        if len(f) < 3 or f[0] == b"":
            continue
        s, e = _stoi(f[1]), _stoi(f[2])
        if s is None or e is None:
            continue
        if normalize and s > e:
            s, e = e, s
        e += end_shift
        if not (I32.min <= e <= I32.max):
            continue
        name = f[0].decode()
        if name not in ids:
            ids[name] = len(names)
            names.append(name)
        contig.append(ids[name]); starts.append(s); ends.append(e)
    return (names, np.array(contig, np.int32), np.array(starts, np.int32), np.array(ends, np.int32), len(lines),
            len(lines) - len(starts))


# ---- the reference's own loader (oracle/_ref/libsi_bedref.so: test/bench.cpp compiled in place) ------------------
def _bedref_path():
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libsi_bedref.so")


def reference_available():
    import os
    return os.path.exists(_bedref_path())


def reference_load(intervals_text: bytes, queries_text: bytes = b"", return_seconds=False):
    """Bench::load_intervals (reference test/bench.cpp:67-102) on the two BED texts, through temporary files:
    ((starts, ends), (q_starts, q_ends)) of the "chr1" records, min/max-normalised, in file order. The loader
    calls std::stoi unguarded: feed it well-formed lines only (a malformed line terminates the process, as it
    terminates the reference's benchmark)."""
    import ctypes as C
    import os
    import tempfile
    L = C.CDLL(_bedref_path())
    L.si_ref_load_intervals.restype = C.c_size_t
    L.si_ref_load_intervals.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.POINTER(C.c_size_t)]
    cap_a = intervals_text.count(b"\n") + 1
    cap_b = queries_text.count(b"\n") + 1
    a_s, a_e = np.zeros(cap_a, np.int32), np.zeros(cap_a, np.int32)
    b_s, b_e = np.zeros(cap_b, np.int32), np.zeros(cap_b, np.int32)
    with tempfile.TemporaryDirectory() as d:
        pa, pb = os.path.join(d, "a.bed"), os.path.join(d, "b.bed")
        with open(pa, "wb") as f:
            f.write(intervals_text)
        with open(pb, "wb") as f:
            f.write(queries_text)
        nq = C.c_size_t(0)
        import time
        t0 = time.perf_counter()
        na = L.si_ref_load_intervals(pa.encode(), pb.encode(), a_s.ctypes.data, a_e.ctypes.data, cap_a, b_s.ctypes.data,
                                     b_e.ctypes.data, cap_b, C.byref(nq))
        dt = time.perf_counter() - t0
    if return_seconds:
        return (a_s[:na], a_e[:na]), (b_s[:nq.value], b_e[:nq.value]), dt
    return (a_s[:na], a_e[:na]), (b_s[:nq.value], b_e[:nq.value])
