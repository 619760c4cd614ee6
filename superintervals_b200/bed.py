"""BED ingest on the device: text -> (contig id, start, end) columns (csrc/bed.cu, siParseBed).

The step before the query path. The reference's callers tokenise BED one line
at a time on a host thread (reference test/bench.cpp:67-102; examples/bed-intersect-si.rs:63-123
keeps one container per chrom); here the whole buffer is uploaded once and split, parsed and
compacted by kernels. Same field rules: tab-separated chrom, start, end; numbers by std::stoi's
rules; further columns ignored. Lines the reference would choke on are skipped and counted.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from . import _lib

__all__ = ["BedTable", "parse_bed", "split_by_contig"]


@dataclass
class BedTable:
    names: List[str]          # distinct chroms in order of first appearance
    contig: np.ndarray        # int32[n]: index into names
    starts: np.ndarray        # int32[n]
    ends: np.ndarray          # int32[n]
    lines: int                # lines seen
    skipped: int              # lines without chrom + numeric start + numeric end
    contig_offsets: np.ndarray = None   # int64[len(names) + 1] when grouped by contig on the device


def parse_bed(source, normalize: bool = False, end_shift: int = 0, group_by_contig: bool = False) -> BedTable:
    """source: a path, bytes, or a uint8 array holding BED text. normalize swaps start/end where
    start > end (bench.cpp:89); end_shift = -1 stores BED's half-open ends inclusively (bench.cpp:210);
    group_by_contig reorders the records by contig on the device (stable) and fills contig_offsets."""
    if isinstance(source, (bytes, bytearray, memoryview)):
        buf = np.frombuffer(source, np.uint8)
    elif isinstance(source, np.ndarray):
        buf = np.ascontiguousarray(source, np.uint8)
    else:
        buf = np.fromfile(source, np.uint8)
    L = _lib.lib()
    t = _lib.siBedTable()
    rc = L.siParseBed(C.cast(buf.ctypes.data, C.c_char_p), buf.size, int(bool(normalize)), int(end_shift),
                      int(bool(group_by_contig)), C.byref(t))
    _lib.check("siParseBed")
    if rc:
        raise RuntimeError(f"siParseBed failed with CUDA error {rc}")
    n = int(t.n)
    take = lambda p: np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int32)
    out = BedTable([t.names[k].decode() for k in range(int(t.n_contigs))], take(t.contig), take(t.starts), take(t.ends),
                   int(t.lines), int(t.skipped))
    if group_by_contig and t.contig_offsets:
        out.contig_offsets = np.ctypeslib.as_array(t.contig_offsets, shape=(int(t.n_contigs) + 1,)).astype(np.int64)
    L.siBedTableFree(C.byref(t))
    return out


def split_by_contig(table: BedTable) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """{chrom: (starts, ends)} in line order inside each chrom -- the per-contig containers of
    bed-intersect-si.rs:100-123, ready for one index per contig (genome.GenomeIndex)."""
    if table.contig_offsets is not None:          # grouped on the device: slices, no host sort
        b = table.contig_offsets
        return {name: (table.starts[b[k]:b[k + 1]], table.ends[b[k]:b[k + 1]]) for k, name in enumerate(table.names)}
    order = np.argsort(table.contig, kind="stable")
    bounds = np.searchsorted(table.contig[order], np.arange(len(table.names) + 1))
    return {name: (table.starts[order[bounds[k]:bounds[k + 1]]], table.ends[order[bounds[k]:bounds[k + 1]]])
            for k, name in enumerate(table.names)}
