#!/usr/bin/env python
"""bench.py -- overlap queries/sec of the batch count path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

Workload (config.workload): BASELINE.json configs[1] -- 10M read-length intervals
(150 bp-10 kb) x 100M range queries on one 250 Mb axis, count only, synthetic,
seeded (superintervals_b200/workloads.py config2). A "step" is one pass of
count over the whole 100M-query batch. Under N>1 every rank holds the same index
and its own independent 100M-query shard (weak scaling, no data-path collective;
only the per-rank hit totals are gathered over NCCL for the CSR base offsets).

value   queries/s, whole job, queries resident in HBM, shuffled order: the step is one
        launch of the rank-cells count kernel on the batch as it arrives (behind the
        locality partition only when the rank cells outgrow L2).
e2e     same metric through the C-ABI host-buffer call countOverlapsBatch()
        (pinned host buffers; H2D of the queries and D2H of the counts inside).
roofline / cpu_baseline: see DESIGN.md section "Measurement".
Secondary objects of the same line: search_values (C3: CSR on heavy-tailed nested
intervals), bed_ingest (BED text -> columns) and set_algebra (mergeOverlaps /
intersection / difference through the C ABI), each with its own CPU baseline.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "overlap queries/sec (count)"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--intervals", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=100_000_000, help="queries per GPU per step")
    ap.add_argument("--order", default="shuffled", choices=["shuffled", "sorted"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c1"],
                    help="c2: BASELINE configs[1] (the metric's workload); c1: configs[0], the reference's own "
                         "generate_test_intervals.py pair (1M x 1M on chr1; sets --intervals/--queries)")
    ap.add_argument("--cpu-sample", type=int, default=16_000_000, help="queries timed on the host CPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-search-values", action="store_true", help="skip the secondary search_values (C3) measurement")
    ap.add_argument("--algo", default="auto", choices=["auto", "walk", "rank", "cells"], help="count kernel (SI_OPT_COUNT_ALGO)")
    ap.add_argument("--partition", action="store_true", help="cells kernel through the locality partition (SI_OPT_CELLS_DIRECT_BYTES=1)")
    ap.add_argument("--sv-intervals", type=int, default=4_000_000)
    ap.add_argument("--sv-queries", type=int, default=4_000_000)
    ap.add_argument("--bed-lines", type=int, default=20_000_000, help="records of the secondary BED-ingest measurement (0 = skip)")
    ap.add_argument("--setop-intervals", type=int, default=1_000_000, help="intervals per set of the secondary set-algebra measurement (0 = skip)")
    return ap.parse_args()


def workload_name(a):
    if a.workload == "c1":
        return (f"C1: reference test/generate_test_intervals.py pair, {a.intervals/1e6:g}M intervals x {a.queries/1e6:g}M queries "
                f"(~2 kb each) on chr1, count, {a.order} queries")
    return (f"C2: {a.intervals/1e6:g}M read-length intervals (150bp-10kb) x {a.queries/1e6:g}M range queries "
            f"per GPU, 250Mb axis, count only, {a.order} queries")


def make_workload(a, rank, nq):
    """(starts, ends, qs, qe) of the chosen workload; rank selects the query shard under weak scaling."""
    from superintervals_b200 import workloads as W
    if a.workload == "c1":
        s, e, qs, qe = W.config1(a.intervals, seed=rank)
        if rank:   # one index for every rank: rank 0's intervals
            s, e, _, _ = W.config1(a.intervals, seed=0)
        return s, e, qs[:nq], qe[:nq]
    s, e = W.config2_intervals(a.intervals, 2)
    qs, qe = W.config2_queries(nq, 2, shard=rank)
    return s, e, qs, qe


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions (B200_PROFILING.md): an NVML
    polling thread (about every 2 ms; nvidia-smi -lms cannot start fast enough for a 40 ms region),
    nvidia-smi as a fallback. Samples carry host timestamps; stop(t0, t1) keeps those inside."""
    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER_CAP = 0x8, 0x20, 0x40, 0x4

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.run, self.t, self.h, self.nv = gpu_index, [], False, None, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nv, self.h, self.run = nv, h, True
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as ex:   # noqa: BLE001
            self.err = repr(ex)
            self.run = False

    def _poll(self):
        nv, h = self.nv, self.h
        while self.run:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self, windows):
        """windows: [(t0, t1), ...] host-clock intervals of the timed regions."""
        if not self.run:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.run = False
        self.t.join(timeout=1)
        rows = [r for r in self.rows if any(t0 <= r[0] <= t1 for t0, t1 in windows)] or self.rows
        sm = [r[1] for r in rows]
        bits = 0
        for r in rows:
            bits |= r[2]
        reasons = [n for n, m in (("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL),
                                  ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER_CAP)) if bits & m]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": self.max_sm, "power_w_max": max(r[3] for r in rows) if rows else None,
                "samples": len(rows), "reasons": reasons, "source": "NVML polled every ~2 ms inside the timed regions"}


def bind_near_gpu(gpu_index):
    """Multi-GPU runs: keep this rank's threads (and therefore its pinned host buffers, first touch) on
    the CPUs NVML reports as local to its GPU, so that eight ranks do not pull their PCIe traffic
    through one socket's memory. Returns the CPU list used, or None when nothing was changed."""
    try:
        import pynvml as nv
        import torch
        nv.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
        h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = nv.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        use = near & os.sched_getaffinity(0)
        if use and use != os.sched_getaffinity(0):
            os.sched_setaffinity(0, use)
            return sorted(use)
    except Exception:   # noqa: BLE001
        pass
    return None


# ---------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on this host's cores
# ---------------------------------------------------------------------------------------
def cpu_reference(a, starts, ends, qs, qe, steps, warmup, threads):
    """Times oracle/_ref (the UNMODIFIED reference C++) when it was compiled, else the
    plain-C port. Returns dict(value q/s, kind, cores, sample, build_s, found, counts_fn)."""
    from oracle.pyoracle import Oracle, Reference
    n_s = min(a.cpu_sample, qs.size)
    sqs, sqe = qs[:n_s], qe[:n_s]
    if a.order == "sorted":
        o = np.argsort(sqs, kind="stable")
        sqs, sqe = np.ascontiguousarray(sqs[o]), np.ascontiguousarray(sqe[o])
    if Reference.available():
        ref = Reference()
        t_build = ref.time_build(starts, ends)              # add() x N + build(), test/bench.cpp:208-213
        times, found = [], 0
        for i in range(warmup + steps):
            t, found = ref.time_count(sqs, sqe, threads)    # test/bench.cpp:240-242 over `threads` chunks
            if i >= warmup:
                times.append(t)
        t1, _ = ref.time_count(sqs[: max(1, n_s // 8)], sqe[: max(1, n_s // 8)], 1)
        kind = "reference"
        detail = f"si::IntervalMap<int,int>::count, g++ -O3 -march={Reference.kind}"
        single = (max(1, n_s // 8)) / t1
        counts = lambda s_, e_: ref.count_batch(s_, e_, 0, threads)
    else:
        t0 = time.perf_counter(); orc = Oracle(starts, ends); t_build = time.perf_counter() - t0
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter(); c = orc.count_batch(sqs, sqe); dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        found, kind, detail, threads = int(c.sum()), "port", "oracle/si_oracle.c scalar port", 1
        single = n_s / float(np.median(times))
        counts = lambda s_, e_: orc.count_batch(s_, e_)
    med = float(np.median(times))
    return {"value": n_s / med, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"first {n_s} queries of rank 0's batch ({a.order}), index of {starts.size} intervals; "
                      f"median of {steps} passes after {warmup} warm-up; {detail}",
            "single_thread_qps": single, "build_s": t_build, "found": int(found), "ms_per_step": med * 1e3,
            "_counts": counts, "_sample": (sqs, sqe)}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0   # N>1: rank 0 alone runs the CPU arm
    starts, ends, qs, qe = make_workload(a, 0, min(a.queries, a.cpu_sample))
    threads = host_threads()
    r = cpu_reference(a, starts, ends, qs, qe, a.steps, a.warmup, threads)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(a), "step": f"one pass of count over a bounded {min(a.queries, a.cpu_sample)}-query sample"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_qps", "build_s")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def kernel_table(recs, steps, nq, passes, a_walk):
    """Group the library's per-launch CUDA-event records: per kernel class the launches per step,
    the mean launch duration and the algorithmic bytes one launch moves."""
    alg = {"pt_histogram": 4.0 * nq,                                             # one read of the starts
           "pt_onesweep": (20.0 + 24.0 * max(0, passes - 1)) / max(1, passes) * nq,  # 8+12 first pass, 12+12 after
           "count_rank": 16.0 * nq,                                              # 12 B record in, 4 B count out
           "count_cells": 76.0 * nq,                                             # 8 B query in, 4 B count out, 2 x 32 B rank-cell sectors
           "count_walk": a_walk * nq}                                            # SURVEY 8d A_count: the reference walk
    out = {}
    for name, ms in recs:
        d = out.setdefault(name, {"launches": 0, "ms_total": 0.0})
        d["launches"] += 1
        d["ms_total"] += ms
    for name, d in out.items():
        d["ms_per_launch"] = d["ms_total"] / d["launches"]
        d["ms_per_step"] = d["ms_total"] / steps
        d["launches_per_step"] = d["launches"] / steps
        if name in alg:
            d["algorithmic_bytes_per_launch"] = alg[name]
            d["achieved_gbs"] = alg[name] / (d["ms_per_launch"] * 1e-3) / 1e9
        del d["ms_total"]
    return out


def bench_search_values(a, torch, L, _lib, rank):
    """Secondary line: search_values (CSR) on BASELINE configs[2] -- heavy-tailed nested intervals."""
    from superintervals_b200 import workloads as W
    from superintervals_b200.device import DeviceIndex, OPT_TIMING, ORDER_UNSORTED
    n, nq = a.sv_intervals, a.sv_queries
    s, e, qs, qe = W.config3(n, nq, 42)
    ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    off, vals = ix.search_values(dqs, dqe, order=ORDER_UNSORTED)
    total = int(off[nq].item())
    counts = torch.empty(nq, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ix.search_values(dqs, dqe, order=ORDER_UNSORTED, counts=counts, offsets=off, out=vals)
    torch.cuda.synchronize()
    ix.set_option(OPT_TIMING, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ix.search_values(dqs, dqe, order=ORDER_UNSORTED, counts=counts, offsets=off, out=vals)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    recs = ix.read_timings()
    ix.set_option(OPT_TIMING, 0)
    per = {}
    for name, t in recs:
        per[name] = per.get(name, 0.0) + t / a.steps
    # e2e through the C ABI with host buffers (searchValuesBatch: H2D queries, D2H offsets + values).
    # The caller owns one cIndexResult and reuses it (clearIndexResult keeps its capacity), as a C
    # caller of the reference would (c_superintervals.h:98-104, 1066-1075); buffers are plain pageable memory.
    import ctypes as C_
    si = L.createSuperIntervals()
    L.siSetHostMirror(si, False)
    L.addIntervals(si, s.ctypes.data, e.ctypes.data, None, s.size)
    L.indexSuperIntervals(si)
    _lib.check("indexSuperIntervals")
    found = L.createIndexResult()
    h_off = np.zeros(nq + 1, np.uint64)
    def e2e_step():
        L.clearIndexResult(C_.byref(found))
        L.searchValuesBatch(si, qs.ctypes.data, qe.ctypes.data, nq, h_off.ctypes.data, C_.byref(found))
    t0 = time.perf_counter(); e2e_step(); first_s = time.perf_counter() - t0     # includes growing the result buffer
    _lib.check("searchValuesBatch")
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / a.e2e_steps
    v2 = np.ctypeslib.as_array(C_.cast(found.data, C_.POINTER(C_.c_int32)), shape=(int(found.size),))
    ok = bool(int(found.size) == total and np.array_equal(h_off.astype(np.int64), off.cpu().numpy())
              and np.array_equal(v2, vals.cpu().numpy()))
    L.destroyIndexResult(C_.byref(found))
    L.destroySuperIntervals(si)
    fill_ms = per.get("fill_runs", per.get("fill", 0.0))
    compulsory = 24.0 * nq + 8.0 * total            # queries + the offset pair in, value read + value written per hit
    peak, _ = measured_peak()
    sv_roof = {"bound": "hbm", "kernel": "fill_runs" if "fill_runs" in per else "fill", "kernel_ms": fill_ms,
               "compulsory_bytes_per_launch": compulsory,
               "achieved": compulsory / (fill_ms * 1e-3) / 1e9 if fill_ms else None, "peak": peak, "unit": "GB/s",
               "frac": (compulsory / (fill_ms * 1e-3) / 1e9 / peak) if fill_ms else None,
               "stab_lists": ix.stab_info()}
    out = {"workload": f"C3: {n/1e6:g}M heavy-tailed nested intervals (Pareto 1.1, <=1Mb) x {nq/1e6:g}M queries, "
                       f"search_values CSR, shuffled queries", "value": nq / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms, "hits": total, "hits_per_query": total / nq, "kernel_ms_per_step": per,
           "result_gbs": (total * 4 + nq * 8) / (ms * 1e-3) / 1e9, "roofline": sv_roof,
           "e2e": {"value": nq / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": 8 * nq,
                   "d2h_bytes_per_step": 8 * (nq + 1) + 4 * total, "equals_device": ok,
                   "first_call_ms": first_s * 1e3,
                   "call": "searchValuesBatch(si, qs, qe, n, offsets, cIndexResult*) with pageable host buffers, result struct reused"}}
    if rank == 0 and not a.no_cpu_baseline:
        from oracle.pyoracle import Reference
        if Reference.available():
            ref = Reference(s, e)
            th = host_threads()
            t, found = ref.time_search_values(qs, qe, th)
            t, found = ref.time_search_values(qs, qe, th)
            out["cpu_baseline"] = {"value": nq / t, "unit": UNIT, "cores": th, "kind": "reference",
                                   "sample": f"all {nq} queries, si::IntervalMap<int,int>::search_values into per-thread vectors",
                                   "found": int(found), "found_equals_gpu": int(found) == total}
    return out


def bench_bed_ingest(a, rank):
    """Secondary line: the step before the path -- BED text to columns (siParseBed) -- beside the
    reference's loader loop (test/bench.cpp:67-102, restated in oracle/bed_cpu.cpp) on one host thread."""
    from superintervals_b200 import workloads as W
    from superintervals_b200.bed import parse_bed
    n = a.bed_lines
    text, cid, s, e = W.bed_text(n)
    parse_bed(text[: 26 * 1000])
    dt = float("inf")
    for _ in range(2):   # best of two: the first call also pays for first-touch of fresh host pages
        t0 = time.perf_counter(); t = parse_bed(text, True, -1); dt = min(dt, time.perf_counter() - t0)
    ok = bool(len(t.starts) == n and np.array_equal(t.starts, s.astype(np.int32)) and np.array_equal(t.ends, (e - 1).astype(np.int32)))
    dg = float("inf")
    for _ in range(2):
        t0 = time.perf_counter(); g = parse_bed(text, True, -1, group_by_contig=True); dg = min(dg, time.perf_counter() - t0)
    ok_g = bool(g.contig_offsets is not None and int(g.contig_offsets[-1]) == n and
                np.array_equal(g.starts[:int(g.contig_offsets[1])], t.starts[t.contig == 0]))   # stable: line order inside a contig
    out = {"workload": f"{n/1e6:g}M BED records ({text.size/1e6:.0f} MB of text, 24 contigs) -> contig/start/end columns",
           "value": n / dt, "unit": "lines/s", "seconds": dt, "text_gb_per_s": text.size / dt / 1e9, "equals_generator": ok,
           "grouped_by_contig": {"seconds": dg, "value": n / dg, "unit": "lines/s", "first_contig_equals_line_order_filter": ok_g},
           "note": "host text buffer in, host columns out: H2D of the text and D2H of the columns are inside; best of 2 calls"}
    so = os.path.join(ROOT, "oracle", "libsi_bedcpu.so")
    if rank == 0 and not a.no_cpu_baseline and os.path.exists(so):
        L = C.CDLL(so)
        L.si_bed_parse_cpu.restype = C.c_size_t
        L.si_bed_parse_cpu.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        m = min(n, 4_000_000)
        cs, ce, cc = (np.empty(m, np.int32) for _ in range(3))
        t0 = time.perf_counter(); got = L.si_bed_parse_cpu(text.ctypes.data, 26 * m, cs.ctypes.data, ce.ctypes.data, cc.ctypes.data, m); dc = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": got / dc, "unit": "lines/s", "cores": 1, "kind": "port",
                               "sample": f"first {m} lines; getline + istringstream + stoi loop of test/bench.cpp:67-102 (oracle/bed_cpu.cpp)",
                               "matches_device": bool(np.array_equal(cs, t.starts[:m]) and np.array_equal(ce - 1, t.ends[:m]))}
    return out


def bench_set_algebra(a, L, _lib, rank):
    """Secondary line: the set-algebra callers of the path through the C ABI (host arrays in, new host
    set out) beside the reference's own sequential C functions (oracle/_ref/libsi_cref.so, one thread)."""
    from superintervals_b200 import workloads as W
    n = a.setop_intervals
    A = W.config2_intervals(n, 11)
    B = W.config2_intervals(n, 12)

    def run(lib, timed):
        def make(s, e, index):
            si = lib.createSuperIntervals()
            if hasattr(lib, "addIntervals"):
                lib.addIntervals(si, s.ctypes.data, e.ctypes.data, None, s.size)
            else:
                for k in range(s.size):
                    lib.addInterval(si, int(s[k]), int(e[k]), k)
            if index:
                lib.indexSuperIntervals(si)
            return si
        a_, b_ = make(*A, False), make(*B, True)
        out = {}
        for name, call in (("mergeOverlaps", lambda: lib.mergeOverlaps(a_, None)), ("intersection", lambda: lib.intersection(a_, b_, None)),
                           ("difference", lambda: lib.difference(a_, b_))):
            best, size, chk = float("inf"), 0, 0
            for _ in range(timed):
                t0 = time.perf_counter(); r = call(); dt = time.perf_counter() - t0
                size = int(r.contents.size)
                if size:
                    st = np.ctypeslib.as_array(r.contents.starts, shape=(size,)); en = np.ctypeslib.as_array(r.contents.ends, shape=(size,))
                    chk = int(st.astype(np.int64).sum() * 3 + en.astype(np.int64).sum())
                lib.destroySuperIntervals(r)
                best = min(best, dt)
            out[name] = {"seconds": best, "pieces": size, "checksum": chk}
        lib.destroySuperIntervals(a_); lib.destroySuperIntervals(b_)
        return out

    ours = run(L, 3)
    _lib.check("set algebra")
    res = {"workload": f"A, B = {n/1e6:g}M read-length intervals each on the 250 Mb axis (B indexed); host arrays in, new host set out",
           "ops": {k: {"seconds": v["seconds"], "pieces": v["pieces"], "input_intervals_per_s": n / v["seconds"]} for k, v in ours.items()}}
    cref = os.path.join(ROOT, "oracle", "_ref", "libsi_cref.so")
    if rank == 0 and not a.no_cpu_baseline and os.path.exists(cref):
        ref = run(_lib.bind(C.CDLL(cref)), 1)
        res["cpu_baseline"] = {"kind": "reference", "cores": 1, "sample": "same sets, reference c_superintervals.h compiled -O3 (oracle/_ref/libsi_cref.so), one pass",
                               "ops": {k: {"seconds": v["seconds"], "pieces": v["pieces"], "equal_to_device": v["pieces"] == ours[k]["pieces"] and v["checksum"] == ours[k]["checksum"]}
                                       for k, v in ref.items()}}
    return res


def main():
    a = parse()
    if a.workload == "c1":
        a.intervals = a.queries = min(a.intervals, 1_000_000) if a.intervals != 10_000_000 else 1_000_000
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist
    from superintervals_b200 import _lib, workloads as W
    from superintervals_b200.device import DeviceIndex, OPT_TIMING, ORDER_SORTED, ORDER_UNSORTED

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: superintervals_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    near_cpus = bind_near_gpu(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()

    # ---- synthetic inputs (seeded): same index on every rank, one query shard per rank
    starts, ends, qs, qe = make_workload(a, rank, a.queries)
    if a.order == "sorted":
        o = np.argsort(qs, kind="stable")
        qs, qe = np.ascontiguousarray(qs[o]), np.ascontiguousarray(qe[o])
    nq = qs.size
    order = ORDER_SORTED if a.order == "sorted" else ORDER_UNSORTED

    ix = DeviceIndex()
    if a.algo != "auto":
        from superintervals_b200.device import COUNT_CELLS, COUNT_RANK, COUNT_WALK, OPT_COUNT_ALGO
        ix.set_option(OPT_COUNT_ALGO, {"walk": COUNT_WALK, "rank": COUNT_RANK, "cells": COUNT_CELLS}[a.algo])
    if a.partition:
        ix.set_option(5, 1)
    for env, opt in (("SIB_BUCKET", 1), ("SIB_WSHIFT", 2), ("SIB_GRID", 4), ("SIB_CELLS_FILL", 6)):      # tuning experiments (tools/variants.sh)
        if os.environ.get(env):
            ix.set_option(opt, int(os.environ[env]))
    d_s, d_e = torch.from_numpy(starts).cuda(), torch.from_numpy(ends).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ix.build(d_s, d_e); torch.cuda.synchronize(); build_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); ix.build(d_s, d_e); torch.cuda.synchronize(); build_ms = min(build_ms, (time.perf_counter() - t0) * 1e3)
    cells_info = ix.cells_info()
    d_qs, d_qe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    counts = torch.empty(nq, dtype=torch.int32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        # shuffled batch: partition by position on the device (records of start, end, index), then count
        ix.count(d_qs, d_qe, out=counts, order=order)

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, device timing, clocks sampled meanwhile
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    ix.set_option(OPT_TIMING, 1)            # CUDA event pair around every hot kernel launch, on its own stream
    launches0 = L.si_b200_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    ev0.record()
    for k in range(a.steps):
        step()
    ev1.record()
    barrier()
    w1 = time.perf_counter()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = int(L.si_b200_kernel_launches() - launches0)
    recs = ix.read_timings()
    ix.set_option(OPT_TIMING, 0)

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t[0])
    ms_per_step = elapsed_ms / a.steps
    value = world * nq / (ms_per_step * 1e-3)

    # ---- per-rank hit totals -> CSR base offsets of each shard (the only exchange: NCCL all_gather)
    hits = counts.to(torch.int64).sum().reshape(1)
    if world > 1:
        allh = [torch.zeros_like(hits) for _ in range(world)]
        dist.all_gather(allh, hits)
        shard_hits = [int(x.item()) for x in allh]
    else:
        shard_hits = [int(hits.item())]
    shard_base = [int(x) for x in np.concatenate([[0], np.cumsum(shard_hits)[:-1]])]

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (pinned), copies inside the timed region
    si = L.createSuperIntervals()
    L.siSetHostMirror(si, False)                        # queries only: skip the index read-back
    L.addIntervals(si, starts.ctypes.data, ends.ctypes.data, None, starts.size)
    L.indexSuperIntervals(si)
    _lib.check("indexSuperIntervals")
    h_qs, h_qe = torch.from_numpy(qs).pin_memory(), torch.from_numpy(qe).pin_memory()
    h_out = torch.empty(nq, dtype=torch.int64).pin_memory()
    def e2e_step():
        L.countOverlapsBatch(si, h_qs.data_ptr(), h_qe.data_ptr(), nq, h_out.data_ptr())
    e2e_step()
    _lib.check("countOverlapsBatch")
    barrier()
    w2 = time.perf_counter()
    for _ in range(a.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    w3 = time.perf_counter()
    e2e_s = (w3 - w2) / a.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    clocks = sampler.stop([(w0, w1), (w2, w3)]) if rank == 0 else None
    e2e_ok = bool((h_out.numpy().astype(np.int64) == counts.cpu().numpy().astype(np.uint32).astype(np.int64)).all())
    L.destroySuperIntervals(si)

    sv = None
    if world == 1 and not a.no_search_values:
        sv = bench_search_values(a, torch, L, _lib, rank)

    bed = None
    if world == 1 and not a.no_search_values and a.bed_lines > 0:
        bed = bench_bed_ingest(a, rank)
    setops = None
    if world == 1 and not a.no_search_values and a.setop_intervals > 0:
        setops = bench_set_algebra(a, L, _lib, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: algorithmic bytes per launch / its own CUDA-event duration
    from oracle.pyoracle import Oracle   # checker only: walk statistics + parity of a sample
    Lg = max(1, math.ceil(math.log2(max(2, a.intervals))))
    stat_n = min(nq, 200_000)
    orc = Oracle(starts, ends)
    h_s, j_s = orc.walk_stats(qs[:stat_n], qe[:stat_n])
    h_per_q = shard_hits[0] / nq                         # exact, from the GPU counts
    j_per_q = j_s / stat_n                               # reference-walk failed tests, sample estimate
    a_walk = 8 + 4 + 4 * Lg + 4 * (h_per_q + j_per_q) + 4 * j_per_q      # SURVEY 8d A_count(q)
    passes = int(round(sum(1 for n_, _ in recs if n_ == "pt_onesweep") / a.steps))
    kernels = kernel_table(recs, a.steps, nq, passes, a_walk)
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    peak, peak_src = measured_peak()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    dk = kernels[dom]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dk.get("achieved_gbs"), "peak": peak, "unit": "GB/s",
                "frac": (dk.get("achieved_gbs") or 0.0) / peak, "peak_source": peak_src, "traffic": traffic,
                "kernel_ms": dk["ms_per_launch"], "launches_per_step": dk["launches_per_step"],
                "share_of_step": dk["ms_per_step"] / ms_per_step,
                "algorithmic_bytes_per_launch": dk.get("algorithmic_bytes_per_launch"),
                "whole_step": {"compulsory_hbm_bytes": 12 * nq + 16 * a.intervals,
                               "compulsory_gbs": (12 * nq + 16 * a.intervals) / (ms_per_step * 1e-3) / 1e9,
                               "reference_walk_bytes_per_query": a_walk, "hits_per_query": h_per_q,
                               "jumps_per_query_sampled": j_per_q, "log2_n": Lg,
                               "reference_walk_gbs": a_walk * nq / (ms_per_step * 1e-3) / 1e9},
                "note": "per-launch CUDA events recorded by the library on the launching stream (SI_OPT_TIMING); "
                        "algorithmic bytes per kernel are defined in DESIGN.md section 4"}

    parity = {"sample": int(stat_n), "mismatches": int((orc.count_batch(qs[:stat_n], qe[:stat_n]).astype(np.int64)
                                                        != counts[:stat_n].cpu().numpy().astype(np.uint32).astype(np.int64)).sum()),
              "e2e_equals_device": e2e_ok}
    del orc

    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_reference(a, starts, ends, qs, qe, 3, 1, host_threads())
        sq, se = r["_sample"]
        m = min(sq.size, 2_000_000)
        if a.order == "shuffled":
            same = int((r["_counts"](sq[:m], se[:m]).astype(np.int64)
                        != counts[:m].cpu().numpy().astype(np.uint32).astype(np.int64)).sum())
            parity["reference_sample"] = m
            parity["reference_mismatches"] = same
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_qps", "build_s")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(a), "intervals": a.intervals, "queries_per_gpu": nq,
                       "step": ((f"device partition of the batch by position ({passes} onesweep passes over 12-byte records) + count kernel"
                                 if passes else "count kernel straight on the shuffled batch (rank cells resident in L2, no partition)")
                                if a.order == "shuffled" else "count kernel on position-sorted queries"),
                       "count_kernel": next((k for k in ("count_cells", "count_rank", "count_walk") if k in kernels), None),
                       "rank_cells": cells_info,
                       "count_algo": a.algo,
                       "l2": (f"inputs larger than L2: {8 * nq / 1e6:.0f} MB of queries + {4 * nq / 1e6:.0f} MB of counts per step vs 126 MB"
                              if 12 * nq > 2 * 126e6 else
                              f"inputs ({12 * nq / 1e6:.0f} MB per step) FIT in L2: a parity/side configuration, not the metric's workload"),
                       "index": "replicated per GPU", "collective": "all_gather of per-rank hit totals (CSR bases)"},
            "e2e": {"value": world * nq / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * nq,
                    "d2h_bytes_per_step": 8 * nq, "ms_per_step": e2e_s * 1e3,
                    "pcie_gbs_each_way": 8 * nq / e2e_s / 1e9,
                    "call": "countOverlapsBatch(si, qs, qe, n, size_t* counts) with pinned host buffers",
                    "rank0_cpus_near_gpu": (len(near_cpus) if near_cpus else None)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "cpu_baseline": cpu_baseline, "parity": parity, "build_ms": build_ms,
            "shard_hits": shard_hits, "shard_csr_base": shard_base, "device_bytes": ix.device_bytes,
            "search_values": sv, "bed_ingest": bed, "set_algebra": setops}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
