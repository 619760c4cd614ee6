"""Pair cells against the separate rank cells: stabbing and range queries on indexes larger than L2.
usage: pair_probe.py [n_intervals] [n_queries]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superintervals_b200.device import DeviceIndex, ORDER_UNSORTED
from superintervals_b200._lib import OPT_PAIR_CELLS


def timed(fn, steps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def gen(n, axis, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    st = (torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64)
    ln = (150 + torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * (10_000 - 150)).to(torch.int64)
    return st.to(torch.int32), torch.clamp(st + ln, max=2**31 - 1).to(torch.int32), g


n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 128_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 256_000_000
for law, axis in (("c5 density (0.5/bp)", 2 * n), ("c4 density (0.032/bp)", min(31 * n, 2_000_000_000))):
    s, e, g = gen(n, axis, 5)
    q = (torch.rand(nq, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64).to(torch.int32)
    # C2's query lengths: log-uniform [1, 10 kb]
    ln = torch.exp(torch.rand(nq, generator=g, device="cuda", dtype=torch.float64) * 9.2103).to(torch.int64).to(torch.int32)
    qe = torch.clamp(q.to(torch.int64) + ln, max=2**31 - 1).to(torch.int32)
    out = {}
    keep = {}
    for mode in (0, 2):
        ix = DeviceIndex(); ix.set_option(OPT_PAIR_CELLS, mode); ix.build(s, e)
        c1 = torch.empty(nq, dtype=torch.int32, device="cuda"); c2 = torch.empty_like(c1)
        out[f"mode{mode}"] = {"stab_ms": timed(lambda: ix.count(q, q, out=c1, order=ORDER_UNSORTED)),
                              "range_ms": timed(lambda: ix.count(q, qe, out=c2, order=ORDER_UNSORTED)),
                              "cells": ix.cells_info(), "device_bytes": ix.device_bytes}
        keep[mode] = (c1, c2)
        del ix
    out["equal"] = bool(torch.equal(keep[0][0], keep[2][0]) and torch.equal(keep[0][1], keep[2][1]))
    print(json.dumps({"law": law, "n": n, "nq": nq, **out}))
    del s, e, q, qe, ln, keep
    torch.cuda.empty_cache()
