"""Driver for the launch list of the rows either side of the path: set algebra through the C ABI
(A, B = 1 M read-length intervals) and BED ingest (5 M lines, grouped by contig).
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_rows.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from superintervals_b200 import _lib, workloads as W
from superintervals_b200.bed import parse_bed

L = _lib.lib()
A = W.config2_intervals(1_000_000, 11)
B = W.config2_intervals(1_000_000, 12)
def make(s, e, index):
    si = L.createSuperIntervals()
    L.addIntervals(si, s.ctypes.data, e.ctypes.data, None, s.size)
    if index:
        L.indexSuperIntervals(si)
    return si
a, b = make(*A, False), make(*B, True)
for call in (lambda: L.mergeOverlaps(a, None), lambda: L.intersection(a, b, None), lambda: L.difference(a, b),
             lambda: L.uniqueIntervals(a, None), lambda: L.intervalGaps(a, 0, 250_000_000, 0), lambda: L.expandIntervals(a, 10, 10, 0, 250_000_000)):
    r = call(); print(int(r.contents.size)); L.destroySuperIntervals(r)
text, *_ = W.bed_text(5_000_000)
t = parse_bed(text, True, -1, group_by_contig=True)
print(len(t.starts), len(t.names))
_lib.check("prof_rows")
