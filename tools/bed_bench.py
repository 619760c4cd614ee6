"""BED ingest at scale: device tokeniser (siParseBed) beside the reference's loader loop on one host
thread (oracle/bed_cpu.cpp <- reference test/bench.cpp:67-102). Text is synthesised in memory
(fixed-width records, 24 contigs); file I/O is excluded on both sides.
usage: python tools/bed_bench.py [lines] [cpu_lines]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superintervals_b200.bed import parse_bed


def synth(n, seed=1):
    """n lines 'chrNN\\tSSSSSSSSS\\tEEEEEEEEE\\n' (zero-padded numbers: std::stoi reads them the same)."""
    rng = np.random.default_rng(seed)
    cid = rng.integers(1, 25, n)
    s = rng.integers(0, 249_000_000, n)
    e = s + rng.integers(1, 10_000, n)
    buf = np.empty((n, 27), np.uint8)
    buf[:, 0:3] = np.frombuffer(b"chr", np.uint8)
    buf[:, 3] = 48 + cid // 10
    buf[:, 4] = 48 + cid % 10
    buf[:, 5] = 9
    for k in range(9):
        buf[:, 6 + k] = 48 + (s // 10 ** (8 - k)) % 10
        buf[:, 16 + k] = 48 + (e // 10 ** (8 - k)) % 10
    buf[:, 15] = 9
    buf[:, 25] = 10
    return buf[:, :26].copy().reshape(-1), cid, s, e          # 26 bytes per line


n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
text, cid, s, e = synth(n)
parse_bed(text[: 26 * 1000])                                      # warm-up: context, kernels
t0 = time.perf_counter(); t = parse_bed(text, True, -1); dt = time.perf_counter() - t0
ok = (len(t.starts) == n and np.array_equal(t.starts, s.astype(np.int32)) and np.array_equal(t.ends, (e - 1).astype(np.int32))
      and [t.names[c] for c in t.contig[:1000]] == [f"chr{c:02d}" for c in cid[:1000]])
L = C.CDLL(os.path.join(ROOT, "oracle", "libsi_bedcpu.so"))
L.si_bed_parse_cpu.restype = C.c_size_t
L.si_bed_parse_cpu.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
cs, ce, cc = (np.empty(n_cpu, np.int32) for _ in range(3))
t0 = time.perf_counter(); m = L.si_bed_parse_cpu(text.ctypes.data, 26 * n_cpu, cs.ctypes.data, ce.ctypes.data, cc.ctypes.data, n_cpu); dc = time.perf_counter() - t0
print(json.dumps({"lines": n, "text_bytes": int(text.size), "device_parse_s": dt, "device_lines_per_s": n / dt,
                  "device_gb_per_s": text.size / dt / 1e9, "equals_generator": bool(ok), "contigs": len(t.names),
                  "cpu_lines": int(m), "cpu_parse_s": dc, "cpu_lines_per_s": m / dc, "cpu_threads": 1,
                  "cpu_matches_device": bool(np.array_equal(cs, t.starts[:n_cpu]) and np.array_equal(ce - 1, t.ends[:n_cpu])),
                  "note": "device time includes H2D of the text (pageable numpy buffer) and D2H of the three columns"}))
