"""BASELINE.json configs[3] (C4: 24 contigs, chromosome-partitioned across GPUs) and configs[4]
(C5: build() throughput on 1 B intervals + 1 B stabbing queries) at up to full size.

    python tools/run_configs.py c5 [--intervals 1000000000 --queries 1000000000]
    python tools/run_configs.py c4 [--scale 1.0]                     # 1 GPU: all 24 contigs on it
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tools/run_configs.py c4 [--scale 1.0]     # N GPUs: contigs by LPT

Inputs are generated ON THE DEVICE (torch CUDA generator, seeded) with the length / position laws
of SURVEY.md 8d, so a billion-row case needs no host memory. Parity at these sizes is checked
through the closed form #{starts <= qe} - #{ends < qs} computed with torch.sort / searchsorted
(every interval is well formed and no query is inverted here). One JSON line per run (rank 0).
Timing: CUDA events on the launching stream, max over ranks.
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from superintervals_b200 import workloads as W
from superintervals_b200.device import DeviceIndex, ORDER_UNSORTED
from superintervals_b200.genome import GenomeIndex

CHUNK = 1 << 26


def log_uniform_len(g, n, lo, hi):
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    for a in range(0, n, CHUNK):
        b = min(n, a + CHUNK)
        u = torch.rand(b - a, generator=g, device="cuda", dtype=torch.float64)
        out[a:b] = torch.exp(u * (math.log(hi) - math.log(lo)) + math.log(lo)).to(torch.int64).clamp_(lo, hi)
    return out


def gen_ranges(g, n, axis, lo, hi):
    """n end-inclusive ranges [s, s + len - 1], len log-uniform[lo, hi], s uniform on [0, axis - len)."""
    ln = log_uniform_len(g, n, lo, hi)
    s = torch.empty(n, dtype=torch.int32, device="cuda")
    e = torch.empty(n, dtype=torch.int32, device="cuda")
    for a in range(0, n, CHUNK):
        b = min(n, a + CHUNK)
        u = torch.rand(b - a, generator=g, device="cuda", dtype=torch.float64)
        st = (u * (axis - ln[a:b]).to(torch.float64)).to(torch.int64)
        s[a:b] = st.to(torch.int32)
        e[a:b] = (st + ln[a:b] - 1).to(torch.int32)
    return s, e


def closed_form(ss, se, qs, qe):
    return (torch.searchsorted(ss, qe, right=True) - torch.searchsorted(se, qs, right=False)).to(torch.int64)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)


def run_c5(a):
    g = torch.Generator(device="cuda").manual_seed(5)
    n, nq, axis = a.intervals, a.queries, 2_000_000_000
    s, e = gen_ranges(g, n, axis, 150, 10_000)
    torch.cuda.synchronize()
    ix = DeviceIndex()
    t0 = time.perf_counter()
    _, build_ms = timed(lambda: ix.build(s, e))
    build_wall = (time.perf_counter() - t0) * 1e3
    dev_bytes = ix.device_bytes
    # stabbing queries in slices (the library slices at 2^27 itself; slicing here bounds the query arrays)
    sl = min(nq, 1 << 27)
    count_ms, hits, mism, checked = 0.0, 0, 0, 0
    ss = se = None
    for at in range(0, nq, sl):
        m = min(sl, nq - at)
        q = (torch.rand(m, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64).to(torch.int32)
        if at == 0:
            ix.count(q, q, order=ORDER_UNSORTED)            # untimed: sizes the scratch buffers (cudaMalloc)
        c, ms = timed(lambda: ix.count(q, q, order=ORDER_UNSORTED))
        count_ms += ms
        hits += int(c.to(torch.int64).sum().item())
        if at == 0:                                   # closed-form parity on the first slice's head
            k = min(m, a.check)
            ss, se = torch.sort(s)[0], torch.sort(e)[0]
            want = closed_form(ss, se, q[:k], q[:k])
            mism = int((c[:k].to(torch.int64) != want).sum().item())
            checked = k
            del ss, se, want
        del q, c
    print(json.dumps({"config": "C5: build() on device + stabbing queries", "intervals": n, "queries": nq,
                      "axis": axis, "build_ms": build_ms, "build_wall_ms": build_wall,
                      "build_intervals_per_s": n / (build_ms * 1e-3), "count_ms": count_ms,
                      "count_queries_per_s": nq / (count_ms * 1e-3), "hits": hits, "hits_per_query": hits / nq,
                      "device_bytes_after_build": dev_bytes, "parity": {"checked": checked, "mismatches": mism,
                                                                          "how": "closed form via torch.sort + searchsorted"}}))


def run_c4(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parts = W.config4_partition(int(100_000_000 * a.scale), int(1_000_000_000 * a.scale), 4)
    names = [f"chr{i}" for i in list(range(1, 23)) + ["X", "Y"]]
    gi = GenomeIndex(names, [p[0] for p in parts], [p[1] for p in parts], rank=rank, world=world)
    build_ms = count_ms = 0.0
    nq_local = n_local = mism = checked = 0
    for c in gi.owned:
        n_c, q_c, L, seed = parts[c]
        g = torch.Generator(device="cuda").manual_seed(seed)
        s, e = gen_ranges(g, n_c, int(L), 150, 10_000)
        _, ms = timed(lambda: gi.build_contig(c, s, e))
        build_ms += ms
        qs, qe = gen_ranges(g, q_c, int(L), 1, 10_000)
        gi.count_contig(c, qs, qe, order=ORDER_UNSORTED)       # untimed: sizes this index's scratch buffers (cudaMalloc)
        cnt, ms = timed(lambda: gi.count_contig(c, qs, qe, order=ORDER_UNSORTED))
        count_ms += ms
        gi.hits[c] = int(cnt.to(torch.int64).sum().item())
        k = min(q_c, a.check)
        want = closed_form(torch.sort(s)[0], torch.sort(e)[0], qs[:k], qe[:k])
        mism += int((cnt[:k].to(torch.int64) != want).sum().item())
        checked += k
        nq_local += q_c; n_local += n_c
        gi._ix.pop(c)                                  # one contig resident at a time
        del s, e, qs, qe, cnt, want
    bases, totals = gi.csr_bases(device="cuda" if world > 1 else None)   # the only exchange: per-contig hit totals
    t = torch.tensor([count_ms, build_ms, float(mism)], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(nq_local), float(n_local), float(checked)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"config": "C4: 24 contigs (GRCh38 lengths), chromosome-partitioned", "n_gpus": world,
                          "scale": a.scale, "intervals": int(tot[1]), "queries": int(tot[0]),
                          "count_ms_max_over_ranks": float(t[0]), "build_ms_max_over_ranks": float(t[1]),
                          "count_queries_per_s": float(tot[0]) / (float(t[0]) * 1e-3),
                          "owner": [int(x) for x in gi.owner], "hits_total": int(totals.sum()),
                          "csr_base_of_last_contig": int(bases[-1]),
                          "parity": {"checked": int(tot[2]), "mismatches_max_over_ranks": int(t[2]),
                                     "how": "closed form via torch.sort + searchsorted, per contig"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--intervals", type=int, default=1_000_000_000)
    ap.add_argument("--queries", type=int, default=1_000_000_000)
    ap.add_argument("--check", type=int, default=4_000_000)
    args = ap.parse_args()
    (run_c5 if args.which == "c5" else run_c4)(args)
