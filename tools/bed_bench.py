"""BED ingest at scale on the device (siParseBed): text synthesised in memory (26-byte records, 24
contigs), parsed into three host columns, checked against the generating columns. The CPU baseline of
the same step (the reference's loader loop on one host thread) is timed by bench.py's cpu_baseline
leg ("bed_ingest" in its JSON line). usage: python tools/bed_bench.py [lines]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superintervals_b200 import workloads as W
from superintervals_b200.bed import parse_bed

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
text, cid, s, e = W.bed_text(n)
parse_bed(text[: 26 * 1000])                                      # warm-up: context, kernels
t0 = time.perf_counter(); t = parse_bed(text, True, -1); dt = time.perf_counter() - t0
ok = (len(t.starts) == n and np.array_equal(t.starts, s.astype(np.int32)) and np.array_equal(t.ends, (e - 1).astype(np.int32))
      and [t.names[c] for c in t.contig[:1000]] == [f"chr{c:02d}" for c in cid[:1000]])
print(json.dumps({"lines": n, "text_bytes": int(text.size), "device_parse_s": dt, "device_lines_per_s": n / dt,
                  "device_gb_per_s": text.size / dt / 1e9, "equals_generator": bool(ok), "contigs": len(t.names),
                  "note": "device time includes H2D of the text (pageable numpy buffer) and D2H of the three columns"}))
