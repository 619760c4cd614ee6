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
