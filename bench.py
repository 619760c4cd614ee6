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
        (pinned host buffers; H2D of the queries and D2H of the counts inside), beside the
        32-bit-count call and the bare-copy ceiling of the same buffers.
roofline  the dominant kernel against (a) its COMPULSORY HBM bytes (12 B/query + its tables
        once) / the measured copy peak = `frac`, (b) ncu's DRAM bytes, (c) for the
        L2-gather kernel its sector rate / the measured L2 gather peak (profiles/l2_peak_r02.json).
sorted  the same batch position-sorted (the reference's fast case): the TMA-staged streaming
        kernel, its roofline, its end-to-end figure and the CPU arm on sorted input.
strong / gather (N > 1): ONE batch cut into N ranges, per-query counts all-gathered over NCCL.
Secondary objects of the same line: search_values (C3), configs (c1, c4 mode B, c5_lite), build,
latency (single-query C calls), bed_ingest and set_algebra, each with its own CPU baseline.
See DESIGN.md section "Measurement".
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
    ap.add_argument("--no-search-values", action="store_true", help="skip every secondary measurement (search_values, configs, build, latency, BED, set algebra)")
    ap.add_argument("--no-sorted", action="store_true", help="skip the position-sorted arm")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs object (c1, c4, c5_lite)")
    ap.add_argument("--c4-scale", type=float, default=1.0, help="C4 size: 1.0 = 100M intervals x 1B queries over 24 contigs")
    ap.add_argument("--c5-intervals", type=int, default=32_000_000)
    ap.add_argument("--c5-full-intervals", type=int, default=1_000_000_000, help="intervals (= stabbing queries) of the full-size C5 arm at N = 1 (0 = skip)")
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
def cpu_reference(a, starts, ends, qs, qe, steps, warmup, threads, order=None, ref=None):
    """Times oracle/_ref (the UNMODIFIED reference C++) when it was compiled, else the
    plain-C port. Returns dict(value q/s, kind, cores, sample, build_s, found, counts_fn)."""
    from oracle.pyoracle import Oracle, Reference
    order = order or a.order
    n_s = min(a.cpu_sample, qs.size)
    sqs, sqe = qs[:n_s], qe[:n_s]
    if order == "sorted":
        o = np.argsort(sqs, kind="stable")
        sqs, sqe = np.ascontiguousarray(sqs[o]), np.ascontiguousarray(sqe[o])
    if Reference.available():
        t_build = None
        if ref is None:
            ref = Reference()
            t_build = ref.time_build(starts, ends)          # add() x N + build(), test/bench.cpp:208-213
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
            "sample": f"first {n_s} queries of rank 0's batch ({order}), index of {starts.size} intervals; "
                      f"median of {steps} passes after {warmup} warm-up; {detail}",
            "single_thread_qps": single, "build_s": t_build, "found": int(found), "ms_per_step": med * 1e3,
            "_counts": counts, "_sample": (sqs, sqe), "_ref": ref}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0   # N>1: rank 0 alone runs the CPU arm
    starts, ends, qs, qe = make_workload(a, 0, min(a.queries, a.cpu_sample))
    threads = host_threads()
    r = cpu_reference(a, starts, ends, qs, qe, a.steps, a.warmup, threads)
    pub = ("value", "unit", "cores", "kind", "sample", "single_thread_qps", "build_s")
    other = "sorted" if a.order == "shuffled" else "shuffled"
    r2 = cpu_reference(a, starts, ends, qs, qe, max(1, a.steps // 2), 1, threads, order=other, ref=r["_ref"])
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(a), "step": f"one pass of count over a bounded {min(a.queries, a.cpu_sample)}-query sample"},
            "cpu_baseline": {k: r[k] for k in pub},
            other: {"value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms_per_step"], "cores": r2["cores"], "sample": r2["sample"],
                    "single_thread_qps": r2["single_thread_qps"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def kernel_table(recs, steps):
    """Group the library's per-launch CUDA-event records: per kernel class the launches per step and the
    mean launch duration."""
    out = {}
    for name, ms in recs:
        d = out.setdefault(name, {"launches": 0, "ms_total": 0.0})
        d["launches"] += 1
        d["ms_total"] += ms
    for name, d in out.items():
        d["ms_per_launch"] = d["ms_total"] / d["launches"]
        d["ms_per_step"] = d["ms_total"] / steps
        d["launches_per_step"] = d["launches"] / steps
        del d["ms_total"]
    return out


def l2_gather_peak(table_bytes):
    """Measured L2 random-sector gather rate (tools/l2_peak.cu, profiles/l2_peak_r02.json) for a table of this size:
    the entry of the smallest measured table that is at least as large. Returns (sectors/s, description) or (None, why)."""
    p = os.path.join(ROOT, "profiles", "l2_peak_r02.json")
    try:
        rows = json.load(open(p))["gather"]
    except Exception as ex:   # noqa: BLE001
        return None, f"profiles/l2_peak_r02.json unreadable: {ex!r}"
    best = {}
    for r in rows:
        mb = r["table_mb"]
        best[mb] = max(best.get(mb, 0.0), r["sectors_per_s_k8"], r["sectors_per_s_k4"], r.get("sectors_per_s_flat", 0.0))
    for mb in sorted(best):
        if mb * 1e6 >= table_bytes * 0.999:
            return best[mb], f"tools/l2_peak.cu on B200 (profiles/l2_peak_r02.json): random 32-byte sector reads from a {mb:g} MB table"
    mb = max(best)
    return best[mb], f"tools/l2_peak.cu on B200 (profiles/l2_peak_r02.json): largest measured table, {mb:g} MB"


def ncu_traffic(kernel):
    """DRAM bytes per launch of a kernel class from the committed ncu capture (profiles/kernel_traffic_r02.json): a
    constant measured under ncu on C2, NOT measured in this run."""
    for name in ("kernel_traffic_r02.json", "kernel_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            try:
                v = json.load(open(tp)).get(kernel, {}).get("dram_bytes_per_launch")
                if v:
                    return float(v), f"profiles/{name} (ncu --set full capture of this kernel on C2; a constant, not measured in this run)"
            except Exception:   # noqa: BLE001
                pass
    return None, None


def count_roofline(kernels, ms_per_step, nq, tables_bytes, table_name, is_c2):
    """Roofline of the dominant count kernel of one step. Algorithmic bytes per launch = its COMPULSORY HBM traffic:
    8 B/query in + 4 B/query out + the kernel's own tables read once."""
    if not kernels:
        return None
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    dk = kernels[dom]
    peak, peak_src = measured_peak()
    per_launch_q = nq / max(1.0, dk["launches_per_step"])
    alg = 12.0 * per_launch_q + float(tables_bytes)
    secs = dk["ms_per_launch"] * 1e-3
    achieved = alg / secs / 1e9
    traffic, traffic_src = ncu_traffic(dom) if is_c2 else (None, None)
    out = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "peak_source": peak_src, "kernel_ms": dk["ms_per_launch"], "launches_per_step": dk["launches_per_step"],
           "share_of_step": dk["ms_per_step"] / ms_per_step,
           "algorithmic_bytes_per_launch": alg,
           "algorithmic_bytes_are": f"compulsory HBM traffic: 8 B/query in + 4 B/query out + {table_name} ({tables_bytes / 1e6:.1f} MB) read once",
           "frac_hbm_compulsory": achieved / peak,
           "traffic": traffic, "traffic_source": traffic_src,
           "frac_hbm_dram": (traffic / secs / 1e9 / peak) if traffic else None,
           "traffic_over_compulsory": (traffic / alg) if traffic else None}
    if dom == "count_cells":
        # what bounds this kernel is the L2 gather: two random 32-byte sectors per query (one rank cell per table),
        # beside the coalesced query/count streams (12 B/query = 0.375 sectors)
        sectors = 2.0 * per_launch_q
        pk, src = l2_gather_peak(tables_bytes)
        out["l2"] = {"gather_sectors_per_launch": sectors, "achieved_sectors_per_s": sectors / secs,
                     "achieved_gbs": sectors * 32 / secs / 1e9, "peak_sectors_per_s": pk, "peak_source": src,
                     "frac_l2": (sectors / secs / pk) if pk else None,
                     "note": "the limiter of the shuffled step: every query gathers one 32-byte rank-cell sector per table from L2"}
        out["limiter"] = "L2 random-sector gather (see l2); HBM is 20-25 % busy"
    elif dom == "count_stream":
        out["limiter"] = "issue slots and load latency of the rank loop (DESIGN.md 4.4c); DRAM traffic equals the compulsory bytes"
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
    # what bounds the fill is the L2 request path, not DRAM: lane-private list scans and one payload gather per hit. Sector
    # requests per launch come from the committed ncu capture of this kernel on this workload (a constant, not measured
    # in this run), the kernel time is this run's; the peak is the measured random-sector read rate of L2 (tools/l2_peak.cu).
    try:
        kt = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic_r02.json"))).get("fill_runs", {})
        if fill_ms and kt.get("sm_sector_requests_to_l2") and n == 4_000_000 and nq == 4_000_000:
            pk, src = l2_gather_peak(32 << 20)
            sv_roof["traffic"] = kt["dram_bytes_per_launch"]
            sv_roof["frac_hbm_dram"] = kt["dram_bytes_per_launch"] / (fill_ms * 1e-3) / 1e9 / peak
            sv_roof["l2"] = {"sm_sector_requests_per_launch": kt["sm_sector_requests_to_l2"],
                             "achieved_sectors_per_s": kt["sm_sector_requests_to_l2"] / (fill_ms * 1e-3), "peak_sectors_per_s": pk,
                             "peak_source": src, "frac_l2": kt["sm_sector_requests_to_l2"] / (fill_ms * 1e-3) / pk if pk else None,
                             "source": "profiles/kernel_traffic_r02.json: global load sectors that missed L1 + store sectors of one launch (ncu)"}
            sv_roof["limiter"] = "L2 sector requests (see l2); DRAM is about 40 % busy"
    except Exception:   # noqa: BLE001
        pass
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
        # the reference's OWN loader (Bench::load_intervals, test/bench.cpp:67-102, compiled in place: oracle/_ref/libsi_bedref.so)
        # keeps "chr1" lines and reads from a file: a one-contig sample, file read (page cache) inside its time
        from oracle import bed_oracle
        if bed_oracle.reference_available():
            rows = np.arange(2_000_000)
            one = ("\n".join(f"chr1\t{int(x)}\t{int(y)}" for x, y in zip(s[rows], e[rows])) + "\n").encode()
            (rs, re_), _, dr = bed_oracle.reference_load(one, return_seconds=True)
            d1 = parse_bed(one, True, 0)
            out["cpu_baseline"]["reference_loader"] = {
                "value": rs.size / dr, "unit": "lines/s", "cores": 1, "kind": "reference",
                "sample": f"{rs.size} one-contig lines through Bench::load_intervals (reads a temporary file from the page cache)",
                "equals_device": bool(np.array_equal(rs, d1.starts) and np.array_equal(re_, d1.ends))}
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


def timed_device(torch, fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_build(a, torch, ix, d_s, d_e, rank):
    """build() on device-resident input: CUDA events around the whole build on its stream (its few host round trips
    -- sortedness flags, table spans -- are inside), SURVEY 8d A_build against the measured copy peak, and
    cub::DeviceRadixSort on the same pairs as a yardstick (tools/cub_sort.cu, test-only binary)."""
    n = d_s.numel()
    best = float("inf")
    for _ in range(3):
        best = min(best, timed_device(torch, lambda: ix.build(d_s, d_e), 1))
    # live 8-bit digits of the 64-bit key (start, ~end): what the onesweep passes actually run
    smax, emax = int(d_s.max().item()), int(d_e.max().item())
    live = lambda v: max(1, (max(1, v).bit_length() + 7) // 8)
    P = live(smax) + live(emax)
    a_build = 12.0 * n + P * 2 * 12.0 * n + 12.0 * n + 4.0 * n + 4.0 * n
    peak, _ = measured_peak()
    out = {"intervals": n, "ms": best, "intervals_per_s": n / (best * 1e-3), "timing": "CUDA events around siIndexBuildDevice on its stream, best of 3, device-resident input",
           "algorithmic_bytes": a_build, "algorithmic_bytes_are": f"SURVEY 8d A_build with P = {P} live 8-bit digits: 12N in + P x 2 x 12N sort passes + 12N gather + 4N ends + 4N branch",
           "achieved_gbs": a_build / (best * 1e-3) / 1e9, "frac_hbm": a_build / (best * 1e-3) / 1e9 / peak,
           "sort_path": {0: "input already sorted: no sort", 1: "narrow: 4 passes over (start, idx) 8-byte pairs + tie fix + 4 keys-only passes over the ends",
                         2: "composite 64-bit key (start, ~end): 8 passes over 12-byte pairs"}.get(ix.last_sort(), "?"),
           "note": "A_build is SURVEY 8d's figure for a composite-key sort; the narrow path moves fewer bytes than that. The build also makes this "
                   "implementation's own tables (all ends sorted, rank cells, rank bits, block-sorted ends, max tree): their bytes are not in A_build"}
    exe = os.path.join(ROOT, "tools", "bin", "cub_sort")
    if rank == 0 and os.path.exists(exe):
        try:
            r = subprocess.run([exe, str(n), "5"], capture_output=True, text=True, timeout=120)
            out["cub_yardstick"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as ex:   # noqa: BLE001
            out["cub_yardstick"] = {"error": repr(ex)}
    return out


def copy_ceiling(torch, dist, world, nq, h_qs, h_qe, h_out, out_bytes_per_query):
    """What the host link allows for this call's buffers: bare cudaMemcpyAsync of the same pinned arrays, H2D (8 B/query)
    and D2H (out_bytes_per_query) on two streams at once, every rank at the same time. Best of 3, host clock."""
    d_a = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_b = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_c = torch.empty(nq * out_bytes_per_query // 4, dtype=torch.int32, device="cuda")
    h_c = h_out.view(torch.int32)[: d_c.numel()]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = float("inf")
    for _ in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_a.copy_(h_qs, non_blocking=True)
            d_b.copy_(h_qe, non_blocking=True)
        with torch.cuda.stream(s2):
            h_c.copy_(d_c, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def bench_strong(a, torch, dist, L, _lib, ix, world, rank, d_qs, d_qe, counts_rank0):
    """Mode A of SURVEY 8e as the north star words it: ONE batch (rank 0's) cut into one contiguous range per GPU, the index
    replicated, each GPU counts its range straight into its slot of the gathered vector, and the per-query counts are
    all-gathered over NCCL (in place: no staging copy); every GPU then scans them into global 64-bit CSR offsets."""
    from superintervals_b200.device import ORDER_UNSORTED
    nq = d_qs.numel()
    gqs, gqe = d_qs.clone(), d_qe.clone()
    dist.broadcast(gqs, 0)
    dist.broadcast(gqe, 0)
    per = (((nq + world - 1) // world) + 7) & ~7          # 32-byte aligned slots
    lo = min(nq, rank * per); hi = min(nq, lo + per); m = hi - lo
    gathered = torch.zeros(world * per, dtype=torch.int32, device="cuda")
    slot = gathered[rank * per:(rank + 1) * per]
    offsets = torch.empty(world * per + 1, dtype=torch.int64, device="cuda")
    mqs, mqe = gqs[lo:hi].contiguous(), gqe[lo:hi].contiguous()

    def count_only():
        if m:
            ix.count(mqs, mqe, out=slot[:m], order=ORDER_UNSORTED)

    def with_gather():
        count_only()
        dist.all_gather_into_tensor(gathered, slot)

    def with_scan():
        with_gather()
        ix.scan(gathered, out=offsets)

    res = {}
    for name, fn in (("count_only", count_only), ("count_gather", with_gather), ("count_gather_scan", with_scan)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        ms = timed_device(torch, fn, a.steps)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t[0])
    # ---- the same step with the collective FUSED into the count kernel: every count is stored into every GPU's copy of the
    # gathered vector as it is produced (peer stores over NVLink, siCountFanoutDevice), then one one-warp kernel exchanges
    # flags (siPeerBarrierDevice). No NCCL call in the step.
    fused = None
    try:
        from superintervals_b200.sharding import PeerGathered
        pg = PeerGathered(per, world, rank)

        def fused_gather():
            arr, sl, ptrs = pg.current()
            if m:
                ix.count_fanout(mqs, mqe, sl[:m], ptrs, order=ORDER_UNSORTED)
            pg.barrier()
            return arr

        def fused_scan():
            ix.scan(fused_gather(), out=offsets)

        fr = {}
        for name, fn in (("count_gather", fused_gather), ("count_gather_scan", fused_scan)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            ms = timed_device(torch, fn, a.steps)
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fr[name] = float(t[0])
        arr = fused_gather()
        torch.cuda.synchronize()
        same = torch.tensor([1.0 if torch.equal(arr, gathered) and not pg.timed_out() else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        fused = {"ms_count_gather": fr["count_gather"], "ms_count_gather_scan": fr["count_gather_scan"],
                 "value": nq / (fr["count_gather"] * 1e-3), "unit": UNIT,
                 "equals_nccl_gather_on_every_gpu": bool(same[0] > 0),
                 "nvlink_bytes_sent_per_gpu_per_step": 4 * per * (world - 1),
                 "how": "count kernel stores each count into every GPU's gathered vector (CUDA IPC peer memory, double-buffered) + "
                        "one-warp flag barrier kernel; no collective library call inside the step"}
        pg.close()
    except Exception as ex:   # noqa: BLE001
        fused = {"unavailable": repr(ex)[:300]}
    ok = None
    if rank == 0:
        got = torch.cat([gathered[r * per: r * per + max(0, min(nq, (r + 1) * per) - r * per)] for r in range(world)])
        ok = bool(torch.equal(got, counts_rank0)) and int(offsets[world * per].item()) == int(counts_rank0.to(torch.int64).sum().item())
    use_fused = bool(fused and fused.get("equals_nccl_gather_on_every_gpu"))
    return {"scaling": "strong", "queries": nq, "queries_per_gpu": per,
            "value": fused["value"] if use_fused else nq / (res["count_gather"] * 1e-3), "unit": UNIT,
            "value_is": ("count fused with the all-gather of the counts (peer stores over NVLink, see fused)" if use_fused
                         else "count, then ncclAllGather of the counts"),
            "ms_count_only": res["count_only"], "ms_count_gather": res["count_gather"], "ms_count_gather_scan": res["count_gather_scan"],
            "value_count_only": nq / (res["count_only"] * 1e-3), "value_nccl": nq / (res["count_gather"] * 1e-3),
            "collective": "ncclAllGather of the per-query uint32 counts, in place (each GPU's count kernel writes its slot of the gathered vector)",
            "nccl_bytes_received_per_gpu_per_step": 4 * per * (world - 1),
            "nccl_bytes_per_step_all_gpus": 4 * per * (world - 1) * world,
            "gather_gbs_per_gpu": (4 * per * (world - 1)) / max(1e-9, (res["count_gather"] - res["count_only"]) * 1e-3) / 1e9,
            "gathered_equals_single_gpu_counts": ok, "fused": fused,
            "timing": "CUDA events on the launching stream (NCCL work joined by torch), max over ranks"}


def bench_search_values_split(a, torch, dist, L, _lib, world, rank):
    """search_values at N > 1 (mode A): the C3 batch cut into N ranges; count -> all-gather of the counts -> scan into GLOBAL
    CSR offsets on every GPU -> each GPU fills its own segment of the global value array (segments stay where they were made)."""
    import ctypes as C_
    from superintervals_b200 import workloads as W
    from superintervals_b200.device import DeviceIndex, ORDER_UNSORTED, FILL_VALUES
    n, nq = a.sv_intervals, a.sv_queries
    s, e, qs, qe = W.config3(n, nq, 42)
    ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    off1, vals1 = ix.search_values(dqs, dqe, order=ORDER_UNSORTED)          # the single-GPU answer, for the check
    per = (((nq + world - 1) // world) + 7) & ~7
    lo = min(nq, rank * per); hi = min(nq, lo + per); m = hi - lo
    gathered = torch.zeros(world * per, dtype=torch.int32, device="cuda")
    slot = gathered[rank * per:(rank + 1) * per]
    offsets = torch.empty(world * per + 1, dtype=torch.int64, device="cuda")
    mqs, mqe = dqs[lo:hi].contiguous(), dqe[lo:hi].contiguous()
    base = int(off1[lo].item()); total = int(off1[hi].item()) - base
    seg = torch.empty(max(1, total), dtype=torch.int32, device="cuda")
    stream = C_.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        if m:
            ix.count(mqs, mqe, out=slot[:m], order=ORDER_UNSORTED)
        dist.all_gather_into_tensor(gathered, slot)
        ix.scan(gathered, out=offsets)
        if m:   # global offsets of my queries; the output pointer is shifted so that offsets[lo] lands on seg[0]
            L.siFillDevice(ix._ix, mqs.data_ptr(), mqe.data_ptr(), m, offsets.data_ptr() + 8 * rank * per, FILL_VALUES,
                           seg.data_ptr() - 4 * base, ORDER_UNSORTED, stream)

    for _ in range(3):
        step()
    _lib.check("siFillDevice (split)")
    torch.cuda.synchronize(); dist.barrier()
    ms = timed_device(torch, step, a.steps)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = torch.tensor([1.0 if (total == 0 or torch.equal(seg[:total], vals1[base:base + total])) else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    # ---- the same step with the gather of the counts fused into the count kernel (peer stores + flag barrier, no NCCL call)
    fused = None
    try:
        from superintervals_b200.sharding import PeerGathered
        pg = PeerGathered(per, world, rank)
        seg.zero_()

        def fused_step():
            arr, sl, ptrs = pg.current()
            if m:
                ix.count_fanout(mqs, mqe, sl[:m], ptrs, order=ORDER_UNSORTED)
            pg.barrier()
            ix.scan(arr, out=offsets)
            if m:
                L.siFillDevice(ix._ix, mqs.data_ptr(), mqe.data_ptr(), m, offsets.data_ptr() + 8 * rank * per, FILL_VALUES,
                               seg.data_ptr() - 4 * base, ORDER_UNSORTED, stream)

        for _ in range(3):
            fused_step()
        _lib.check("siFillDevice (split, fused)")
        torch.cuda.synchronize(); dist.barrier()
        ms_f = timed_device(torch, fused_step, a.steps)
        tf = torch.tensor([ms_f], dtype=torch.float64, device="cuda")
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        okf = torch.tensor([1.0 if ((total == 0 or torch.equal(seg[:total], vals1[base:base + total])) and not pg.timed_out()) else 0.0],
                           dtype=torch.float64, device="cuda")
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        fused = {"value": nq / (float(tf[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(tf[0]), "segments_equal_single_gpu_values": bool(okf[0] > 0.5),
                 "how": "counts fanned out to every GPU's gathered vector by the count kernel (peer stores over NVLink) + flag barrier kernel, then scan and fill"}
        pg.close()
    except Exception as ex:   # noqa: BLE001
        fused = {"unavailable": repr(ex)[:300]}
    use_fused = bool(fused and fused.get("segments_equal_single_gpu_values"))
    return {"workload": f"C3: {n/1e6:g}M heavy-tailed nested intervals x {nq/1e6:g}M queries, search_values CSR, ONE batch cut into {world} ranges",
            "scaling": "strong", "value": fused["value"] if use_fused else nq / (float(t[0]) * 1e-3), "unit": UNIT,
            "ms_per_step": fused["ms_per_step"] if use_fused else float(t[0]),
            "value_is": "count fused with the gather of the counts" if use_fused else "count, ncclAllGather, scan, fill",
            "value_nccl": nq / (float(t[0]) * 1e-3), "ms_per_step_nccl": float(t[0]), "fused": fused,
            "hits": int(off1[nq].item()), "collective": "ncclAllGather of per-query counts -> global 64-bit offsets on every GPU; value segments stay on their GPU",
            "nccl_bytes_received_per_gpu_per_step": 4 * per * (world - 1),
            "segments_equal_single_gpu_values": bool(ok[0] > 0.5)}


def gen_ranges_device(torch, g, n, axis, lo, hi):
    """n end-inclusive ranges on the device: length log-uniform[lo, hi], start uniform on [0, axis - len) (SURVEY 8d laws)."""
    CH = 1 << 26
    s = torch.empty(n, dtype=torch.int32, device="cuda")
    e = torch.empty(n, dtype=torch.int32, device="cuda")
    for at in range(0, n, CH):
        b = min(n, at + CH)
        u = torch.rand(b - at, generator=g, device="cuda", dtype=torch.float64)
        ln = torch.exp(u * (math.log(hi) - math.log(lo)) + math.log(lo)).to(torch.int64).clamp_(lo, hi)
        u = torch.rand(b - at, generator=g, device="cuda", dtype=torch.float64)
        st = (u * (axis - ln).to(torch.float64)).to(torch.int64)
        s[at:b] = st.to(torch.int32)
        e[at:b] = (st + ln - 1).to(torch.int32)
    return s, e


def reference_counts(starts, ends, qs, qe):
    """Counts of a sample from the compiled reference (oracle/_ref) when present, else the oracle port: the checker."""
    from oracle.pyoracle import Oracle, Reference
    if Reference.available():
        return Reference(starts, ends).count_batch(qs, qe, 0, host_threads()).astype(np.int64), "reference"
    return Oracle(starts, ends).count_batch(qs, qe).astype(np.int64), "port"


def bench_c4(a, torch, dist, world, rank):
    """BASELINE configs[3]: 24 contigs (GRCh38 lengths), 100M intervals x 1B queries x --c4-scale, one index per contig,
    contigs owned by ranks (LPT). The queries arrive as ONE MIXED batch (contig id, start, end), each rank holding a
    slice: routed on the device, dispatched to the owners by all-to-all, counted, combined back (genome.py)."""
    from superintervals_b200 import workloads as W
    from superintervals_b200.genome import GenomeIndex
    parts = W.config4_partition(int(100_000_000 * a.c4_scale), int(1_000_000_000 * a.c4_scale), 4)
    names = [f"chr{i}" for i in list(range(1, 23)) + ["X", "Y"]]
    gi = GenomeIndex(names, [p[0] for p in parts], [p[1] for p in parts], rank=rank, world=world)
    build_ms = 0.0
    n_local = 0
    sample = None
    for c in gi.owned:
        n_c, q_c, Lc, seed = parts[c]
        g = torch.Generator(device="cuda").manual_seed(seed)
        s, e = gen_ranges_device(torch, g, n_c, int(Lc), 150, 10_000)
        build_ms += timed_device(torch, lambda: gi.build_contig(c, s, e), 1)
        n_local += n_c
        if c == 20:   # chr21, the smallest: its intervals go to the host for the reference check
            sample = (s.cpu().numpy(), e.cpu().numpy())
        del s, e
    # this rank's slice of the mixed batch: queries of every contig, in proportion, shuffled
    nq_total = sum(p[1] for p in parts)
    g = torch.Generator(device="cuda").manual_seed(4000 + rank)
    mine = [p[1] // world + (1 if rank < p[1] % world else 0) for p in parts]
    m = sum(mine)
    cid = torch.repeat_interleave(torch.arange(24, device="cuda", dtype=torch.int32), torch.tensor(mine, device="cuda"))
    qs = torch.empty(m, dtype=torch.int32, device="cuda"); qe = torch.empty(m, dtype=torch.int32, device="cuda")
    at = 0
    for c, k in enumerate(mine):
        qs[at:at + k], qe[at:at + k] = gen_ranges_device(torch, g, k, int(parts[c][2]), 1, 10_000)
        at += k
    perm = torch.randperm(m, generator=g, device="cuda")
    cid, qs, qe = cid[perm].contiguous(), qs[perm].contiguous(), qe[perm].contiguous()
    del perm
    out = gi.count_mixed(cid, qs, qe)                      # untimed: sizes every scratch buffer
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    ms = timed_device(torch, lambda: gi.count_mixed(cid, qs, qe), 2)
    t = torch.tensor([ms, build_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(m), float(n_local), float(gi.last_exchange["dispatch_bytes"]), float(gi.last_exchange["combine_bytes"])],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    # parity: this rank's chr21 queries against the compiled reference (the owner of chr21 has its intervals on the host)
    sel = (cid == 20).nonzero().flatten()[:200_000]
    par = {"checked": 0, "mismatches": 0, "against": None}
    have = torch.tensor([1.0 if sample is not None else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:   # the owner's rank number, so that it can share the intervals with the checker (rank 0)
        owner20 = int(gi.owner[20])
        if owner20 != 0:
            n20 = parts[20][0]
            buf = torch.empty(2 * n20, dtype=torch.int32, device="cuda")
            if rank == owner20:
                buf[:n20] = torch.from_numpy(sample[0]).cuda(); buf[n20:] = torch.from_numpy(sample[1]).cuda()
            dist.broadcast(buf, owner20)
            if rank == 0:
                sample = (buf[:n20].cpu().numpy(), buf[n20:].cpu().numpy())
    if rank == 0 and sample is not None and sel.numel():
        want, kind = reference_counts(sample[0], sample[1], qs[sel].cpu().numpy(), qe[sel].cpu().numpy())
        got = out[sel].cpu().numpy().astype(np.uint32).astype(np.int64)
        par = {"checked": int(sel.numel()), "mismatches": int((got != want).sum()), "against": f"{kind} (oracle/_ref si::IntervalMap) on chr21's queries of rank 0's slice"}
    hits = int(out.to(torch.int64).sum().item())
    # ---- mode B without the dispatch: the slices stay in peer-mapped memory, every GPU walks all of them in place over NVLink and
    # answers the queries of its own contigs (siCountMixedPeerDevice); two flag barriers through peer memory, no collective
    peer = None
    if world > 1:
        from superintervals_b200.sharding import PeerBatch
        pb = None
        try:
            cap = torch.tensor([float(m)], dtype=torch.float64, device="cuda")
            dist.all_reduce(cap, op=dist.ReduceOp.MAX)
            pb = PeerBatch(int(cap.item()), world, rank)
            pb.contig[:m].copy_(cid); pb.qs[:m].copy_(qs); pb.qe[:m].copy_(qe)
            pb.set_length(m)
            got = gi.count_mixed_peer(pb)
            # [answers equal the dispatched ones, the kernel could run] on EVERY rank: the ranks take the timed branch together or not at all
            flag = torch.tensor([1.0 if (got is not None and torch.equal(got, out)) else 0.0, 1.0 if got is not None else 0.0],
                                dtype=torch.float64, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if flag[1] > 0:
                torch.cuda.synchronize(); dist.barrier()
                ms_p = timed_device(torch, lambda: gi.count_mixed_peer(pb), 3)
                tp = torch.tensor([ms_p], dtype=torch.float64, device="cuda")
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
                away = torch.tensor([float(sum(n for r, n in enumerate(pb.lengths) if r != rank))], dtype=torch.float64, device="cuda")
                dist.all_reduce(away, op=dist.ReduceOp.SUM)
                peer = {"value": float(tot[0]) / (float(tp[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(tp[0]),
                        "equals_dispatched_counts": bool(flag[0] > 0), "timed_out": pb.timed_out(),
                        "nvlink_contig_id_bytes_per_step": int(4 * away.item()),
                        "nvlink_note": "every GPU reads the 4-byte contig id of every remote query; the owner also reads 8 B and stores 4 B per remote query of its contigs",
                        "how": "slices stay in CUDA-IPC peer memory; each GPU walks all slices in place over NVLink, answers its own contigs' queries and "
                               "stores the counts into the slice they belong to (siCountMixedPeerDevice + 2 flag barriers): no all-to-all, no routing sort, no scatter"}
            else:
                peer = {"unsupported": "a rank holds no rank-cell index"}
        except Exception as ex:   # noqa: BLE001
            peer = {"error": repr(ex)}
        finally:
            if pb is not None:
                pb.close()
    # ---- the alternative at N > 1: every GPU holds ALL 24 indexes (3.8 GB for 100 M intervals) and counts its own slice in one
    # launch -- no routing, no exchange. Moving a query to its owner costs 16 B over NVLink, as much time as counting it from
    # HBM-resident rank cells, so partitioning the index pays only when it does not fit on one GPU.
    repl = None
    if world > 1:
        full = GenomeIndex(names, [p[0] for p in parts], [p[1] for p in parts], rank=0, world=1)
        rb = 0.0
        for c in range(24):
            n_c, q_c, Lc, seed = parts[c]
            g = torch.Generator(device="cuda").manual_seed(seed)
            s, e = gen_ranges_device(torch, g, n_c, int(Lc), 150, 10_000)
            rb += timed_device(torch, lambda: full.build_contig(c, s, e), 1)
            del s, e
        out_r = full.count_mixed(cid, qs, qe)
        torch.cuda.synchronize(); dist.barrier()
        ms_r = timed_device(torch, lambda: full.count_mixed(cid, qs, qe), 3)
        tr = torch.tensor([ms_r, rb], dtype=torch.float64, device="cuda")
        same = torch.tensor([1.0 if torch.equal(out_r, out) else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        repl = {"index": "all 24 contig indexes on every GPU", "value": float(tot[0]) / (float(tr[0]) * 1e-3), "unit": UNIT,
                "ms_per_step": float(tr[0]), "build_ms_max_over_ranks": float(tr[1]), "nccl_bytes_per_step": 0,
                "step": "one launch of the mixed-batch count kernel on each GPU's slice, caller's order", "equals_partitioned_counts": bool(same[0] > 0)}
        del full, out_r
    step = ("device routing by owner (radix sort of the slot ids + gather), all-to-all dispatch, ONE mixed-batch count launch on each owner, "
            "all-to-all combine, scatter back to the caller's order") if world > 1 else \
           "ONE launch of the mixed-batch count kernel over the batch in the caller's order (per-contig rank-cell descriptors in shared memory): nothing routed"
    dispatched = {"value": float(tot[0]) / (float(t[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(t[0]), "step": step}
    use_peer = bool(peer and peer.get("equals_dispatched_counts") and not peer.get("timed_out"))
    return {"workload": f"C4: 24 contigs (GRCh38 lengths), {int(tot[1])} intervals x {int(tot[0])} queries as one mixed batch, one index per contig, "
                        f"contigs owned by {world} GPU(s) (LPT)", "scaling": "strong",
            "value": peer["value"] if use_peer else dispatched["value"], "unit": UNIT,
            "ms_per_step": peer["ms_per_step"] if use_peer else dispatched["ms_per_step"], "build_ms_max_over_ranks": float(t[1]),
            "value_is": ("queries read in place over NVLink peer memory, counts stored in place (see peer); equal to the dispatched counts" if use_peer
                         else "routed + all-to-all dispatch / combine" if world > 1 else "one launch"),
            "step": peer["how"] if use_peer else step, "peer": peer, "dispatched": dispatched if world > 1 else None,
            "index": "partitioned by contig (each GPU holds the contigs it owns)" if world > 1 else "all 24 contig indexes on the one GPU",
            "nccl_dispatch_bytes_per_step": int(tot[2]), "nccl_combine_bytes_per_step": int(tot[3]), "hits_rank0": hits, "parity": par,
            "owner": [int(x) for x in gi.owner], "replicated": repl,
            "pair_cells": (lambda own: gi.index(own[0]).cells_info()["pair"] if own else None)(gi.owned),
            "pair_cells_are": "this rank's first contig: both ranks of a coordinate cell in one 32-byte record, one sector per query whose ends share a cell "
                              "(GenomeIndex builds them when the owned contigs' rank cells cannot share L2)"}



def cells_read_bytes(ci):
    """bytes of rank tables a count reads: the pair cells when they answer, else both rank-cell tables"""
    if ci.get("pair", {}).get("format"):
        return ci["pair"]["bytes"]
    return ci["starts"]["bytes"] + ci["ends"]["bytes"]


def hbm_gather_peak(table_bytes):
    """Measured random 32-byte-sector read rate from a table beyond L2 (tools/hbm_gather.cu, profiles/hbm_gather_r02.json), in the
    count kernels' own launch shape (one request per thread, one CTA per 128 requests): the entry of the smallest measured table
    that is at least as large. Returns (sectors/s, description) or (None, why)."""
    try:
        rows = json.load(open(os.path.join(ROOT, "profiles", "hbm_gather_r02.json")))["gather"]
    except Exception as ex:   # noqa: BLE001
        return None, f"profiles/hbm_gather_r02.json unreadable: {ex!r}"
    best = {}
    for r in rows:
        best[r["table_gb"]] = max(best.get(r["table_gb"], 0.0), r.get("single_per_s_flat", 0.0))
    for gb in sorted(best):
        if gb * 2**30 >= table_bytes * 0.999:
            return best[gb], f"tools/hbm_gather.cu on B200 (profiles/hbm_gather_r02.json): random 32-byte sector reads from a {gb:g} GB table, one per thread"
    gb = max(best)
    return best[gb], f"tools/hbm_gather.cu on B200 (profiles/hbm_gather_r02.json): largest measured table, {gb:g} GB"


def sector_gather(ci, nq, ms, stabs):
    """the random-sector roofline of a count whose tables live in HBM: sectors gathered per second against the measured rate"""
    paired = bool(ci.get("pair", {}).get("format"))
    per_query = 1.0 if (paired and stabs) else 2.0
    rate = per_query * nq / (ms * 1e-3)
    peak, src = hbm_gather_peak(cells_read_bytes(ci))
    return {"sectors_per_query": per_query, "achieved_sectors_per_s": rate, "peak_sectors_per_s": peak,
            "frac": rate / peak if peak else None, "peak_source": src,
            "note": "beyond L2 a gather is charged per 32-byte sector wherever the sectors lie (an adjacent pair costs two), so the kernel's "
                    "roofline is this rate, not the copy bandwidth; a table only a little larger than L2 is partly served from it and can exceed the entry used"}


def bench_configs(a, torch, dist, L, _lib, world, rank):
    """The BASELINE configs that are not the headline (C2) nor the search_values object (C3), each with parity against the
    compiled reference on a sample and its own compulsory-HBM fraction."""
    from superintervals_b200 import workloads as W
    from superintervals_b200.device import DeviceIndex, OPT_TIMING, ORDER_UNSORTED, ORDER_SORTED
    peak, _ = measured_peak()
    res = {}
    # ---- C1: the reference's own generator pair, 1M x 1M (test/generate_test_intervals.py:30-41), both orders
    s, e, qs, qe = W.config1(1_000_000, seed=0)
    ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    cnt = torch.empty(qs.size, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ix.count(dqs, dqe, out=cnt, order=ORDER_UNSORTED)
    ms = timed_device(torch, lambda: ix.count(dqs, dqe, out=cnt, order=ORDER_UNSORTED), 20)
    o = torch.argsort(dqs, stable=True)
    sqs, sqe = dqs[o].contiguous(), dqe[o].contiguous()
    cnt_s = torch.empty_like(cnt)
    for _ in range(3):
        ix.count(sqs, sqe, out=cnt_s, order=ORDER_SORTED)
    ms_s = timed_device(torch, lambda: ix.count(sqs, sqe, out=cnt_s, order=ORDER_SORTED), 20)
    off, vals = ix.search_values(dqs, dqe, order=ORDER_UNSORTED)
    ms_sv = timed_device(torch, lambda: ix.search_values(dqs, dqe, order=ORDER_UNSORTED, counts=cnt, offsets=off, out=vals), 10)
    want, kind = reference_counts(s, e, qs, qe)
    got = cnt.cpu().numpy().astype(np.uint32).astype(np.int64)
    from oracle.pyoracle import Reference
    sv_ok = None
    if Reference.available():
        roff, rvals = Reference(s, e).search_values_batch(qs[:200_000], qe[:200_000])
        # exact (start, end) duplicates may come out in either order (quirk Q3): compare the lists as sorted multisets per query
        ours = vals[: int(off[200_000].item())].cpu().numpy()
        sv_ok = bool(np.array_equal(roff.astype(np.int64), off[:200_001].cpu().numpy()))
        if sv_ok:
            seg = np.repeat(np.arange(200_000), np.diff(roff.astype(np.int64)))
            sv_ok = bool(np.array_equal(np.lexsort((ours, seg)).size, ours.size) and
                         np.array_equal(ours[np.lexsort((ours, seg))], rvals[np.lexsort((rvals, seg))]))
    ci = ix.cells_info()
    res["c1"] = {"workload": "C1: reference test/generate_test_intervals.py pair, 1M intervals x 1M queries (~2 kb) on chr1; inputs fit L2",
                 "count": {"value": qs.size / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "order": "shuffled"},
                 "count_sorted": {"value": qs.size / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s,
                                  "equals_shuffled": bool(torch.equal(cnt_s, cnt[o]))},
                 "search_values": {"value": qs.size / (ms_sv * 1e-3), "unit": UNIT, "ms_per_step": ms_sv, "hits": int(off[-1].item()),
                                   "first_200k_lists_equal_reference": sv_ok},
                 "frac_hbm_compulsory": (12.0 * qs.size + ci["starts"]["bytes"] + ci["ends"]["bytes"]) / (ms * 1e-3) / 1e9 / peak,
                 "parity": {"checked": int(qs.size), "mismatches": int((got != want).sum()), "against": kind, "hits": int(want.sum())}}
    del ix, dqs, dqe, cnt, cnt_s, off, vals
    # ---- C5 lite: build() throughput + stabbing queries, same density as C5 (1 B intervals on a 2e9 axis), scaled to fit a bench run
    n5 = a.c5_intervals
    axis = int(2_000_000_000 * (n5 / 1_000_000_000))
    g = torch.Generator(device="cuda").manual_seed(5)
    s5, e5 = gen_ranges_device(torch, g, n5, axis, 150, 10_000)
    ix = DeviceIndex()
    bms = min(timed_device(torch, lambda: ix.build(s5, e5), 1) for _ in range(2))
    q5 = (torch.rand(2 * n5, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64).to(torch.int32)
    c5 = torch.empty(2 * n5, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ix.count(q5, q5, out=c5, order=ORDER_UNSORTED)
    ix.set_option(OPT_TIMING, 1)
    ms5 = timed_device(torch, lambda: ix.count(q5, q5, out=c5, order=ORDER_UNSORTED), 5)
    k5 = kernel_table(ix.read_timings(), 5)
    ix.set_option(OPT_TIMING, 0)
    m = 200_000
    want, kind = reference_counts(s5.cpu().numpy(), e5.cpu().numpy(), q5[:m].cpu().numpy(), q5[:m].cpu().numpy())
    got = c5[:m].cpu().numpy().astype(np.uint32).astype(np.int64)
    ci = ix.cells_info()
    res["c5_lite"] = {"workload": f"C5 scaled: {n5/1e6:g}M intervals (150bp-10kb) on a {axis/1e6:g} Mb axis (C5's density: ~1170 hits per stab), "
                                  f"build() on device + {2*n5/1e6:g}M stabbing queries",
                      "build": {"ms": bms, "intervals_per_s": n5 / (bms * 1e-3)},
                      "count": {"value": 2 * n5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5, "kernels": k5,
                                "hits_per_query": float(c5.to(torch.float64).mean().item())},
                      "rank_cells": ci,
                      "frac_hbm_compulsory": (12.0 * 2 * n5 + cells_read_bytes(ci)) / (ms5 * 1e-3) / 1e9 / peak,
                      "hbm_sector_gather": sector_gather(ci, 2 * n5, ms5, stabs=True),
                      "parity": {"checked": m, "mismatches": int((got != want).sum()), "against": kind}}
    del ix, s5, e5, q5, c5
    torch.cuda.empty_cache()
    # ---- C5 at FULL size (BASELINE configs[4]): 1 B intervals on a 2e9 axis, build() on device, 1 B stabbing queries. One GPU
    # only (at N > 1 every rank would repeat it) and only where the memory is there (about 90 GB at the peak of the build).
    n5f = a.c5_full_intervals
    free_b = torch.cuda.mem_get_info()[0]
    if world == 1 and n5f > 0 and free_b > 110 * (1 << 30) * (n5f / 1e9):
        axis = 2_000_000_000
        g = torch.Generator(device="cuda").manual_seed(5)
        s5, e5 = gen_ranges_device(torch, g, n5f, axis, 150, 10_000)
        ix = DeviceIndex()
        bms = timed_device(torch, lambda: ix.build(s5, e5), 1)
        # the parity sample first, while the inputs are still there: a 20 Mb window of the axis. Every interval that overlaps a
        # point of the window starts within 10 kb before it, so the reference built on those intervals alone answers the
        # window's stabbing queries exactly as one built on all 1 B would.
        w0, w1 = 700_000_000, 720_000_000
        keep = ((s5 >= w0 - 10_000) & (s5 <= w1)).nonzero().flatten()
        hs, he = s5[keep].cpu().numpy(), e5[keep].cpu().numpy()
        del s5, e5, keep
        torch.cuda.empty_cache()
        q5 = torch.empty(n5f, dtype=torch.int32, device="cuda")
        CH = 1 << 27
        for at in range(0, n5f, CH):
            b = min(n5f, at + CH)
            q5[at:b] = (torch.rand(b - at, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64).to(torch.int32)
        c5 = torch.empty(n5f, dtype=torch.int32, device="cuda")
        ix.count(q5, q5, out=c5, order=ORDER_UNSORTED)
        ms5 = timed_device(torch, lambda: ix.count(q5, q5, out=c5, order=ORDER_UNSORTED), 3)
        sel = ((q5 >= w0) & (q5 <= w1)).nonzero().flatten()[:200_000]
        hq = q5[sel].cpu().numpy()
        want, kind = reference_counts(hs, he, hq, hq)
        got = c5[sel].cpu().numpy().astype(np.uint32).astype(np.int64)
        ci = ix.cells_info()
        res["c5"] = {"workload": f"C5 (BASELINE configs[4]): {n5f/1e9:g} B intervals (150bp-10kb) on a 2e9 axis, build() on device + {n5f/1e9:g} B stabbing queries",
                     "build": {"ms": bms, "intervals_per_s": n5f / (bms * 1e-3), "sort_path": ix.last_sort()},
                     "count": {"value": n5f / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5, "hits_per_query": float(c5[:10_000_000].to(torch.float64).mean().item()),
                               "step": ("one launch of the rank-cells kernel on the pair cells: both ranks of a stabbing query in ONE 32-byte sector gathered from HBM"
                                        if ci["pair"]["format"] else
                                        "one launch of the rank-cells kernel: both 32-byte sectors of a query gathered from HBM (the tables are far larger than L2)")},
                     "hbm_sector_gather": sector_gather(ci, n5f, ms5, stabs=True),
                     "rank_cells": ci, "device_bytes": ix.device_bytes,
                     "frac_hbm_compulsory": (12.0 * n5f + cells_read_bytes(ci)) / (ms5 * 1e-3) / 1e9 / peak,
                     "parity": {"checked": int(sel.numel()), "mismatches": int((got != want).sum()), "against": kind,
                                "how": f"the stabbing queries that fall in [{w0}, {w1}] against the reference built on the {hs.size} intervals that can reach that window"}}
        del ix, q5, c5
        torch.cuda.empty_cache()
    else:
        res["c5"] = {"skipped": "full-size C5 runs on one GPU with enough free memory" if n5f > 0 else "--c5-full-intervals 0"}
    return res


def bench_latency(a, L, _lib, starts, ends, qs, qe):
    """The reference's drivers call one query at a time (test/bench.cpp:219-222,240-242): microseconds per C call
    through the mapped-mailbox path (one launch + one stream synchronise, no copies), beside the reference on one host thread."""
    import ctypes as C_
    si = L.createSuperIntervals()
    L.siSetHostMirror(si, False)
    L.addIntervals(si, starts.ctypes.data, ends.ctypes.data, None, starts.size)
    L.indexSuperIntervals(si)
    _lib.check("indexSuperIntervals")
    k = 2000
    q = [(int(qs[i]), int(qe[i])) for i in range(k)]
    L.siIndexSetOption(L.siIndexOf(si), _lib.OPT_RESIDENT_QUERIES, 0)          # first: one kernel launch per call
    for s_, e_ in q[:50]:
        L.countOverlaps(si, s_, e_)
    t0 = time.perf_counter()
    tot = 0
    for s_, e_ in q:
        tot += L.countOverlaps(si, s_, e_)
    t_count = (time.perf_counter() - t0) / k
    r = L.createIndexResult()
    t0 = time.perf_counter()
    tot_v = 0
    for s_, e_ in q:
        L.clearIndexResult(C_.byref(r))
        L.searchValues(si, s_, e_, C_.byref(r))
        tot_v += int(r.size)
    t_search = (time.perf_counter() - t0) / k
    L.destroyIndexResult(C_.byref(r))
    _lib.check("single-query calls")
    # the same loops with the resident polling warp (SI_OPT_RESIDENT_QUERIES): no launch per call
    res_out = None
    if L.siIndexSetOption(L.siIndexOf(si), _lib.OPT_RESIDENT_QUERIES, 1) == 0:
        for s_, e_ in q[:50]:
            L.countOverlaps(si, s_, e_)
        t0 = time.perf_counter()
        tot_r = 0
        for s_, e_ in q:
            tot_r += L.countOverlaps(si, s_, e_)
        t_count_r = (time.perf_counter() - t0) / k
        r = L.createIndexResult()
        t0 = time.perf_counter()
        tot_vr = 0
        for s_, e_ in q:
            L.clearIndexResult(C_.byref(r))
            L.searchValues(si, s_, e_, C_.byref(r))
            tot_vr += int(r.size)
        t_search_r = (time.perf_counter() - t0) / k
        L.destroyIndexResult(C_.byref(r))
        res_out = {"countOverlaps_us": t_count_r * 1e6, "searchValues_us": t_search_r * 1e6, "same_answers": bool(tot_r == tot and tot_vr == tot_v),
                   "path": "SI_OPT_RESIDENT_QUERIES = 1 (the default): one resident warp polls the mapped pinned mailbox (leaves after 0.2 ms idle / 2 ms at most, relaunched on demand)"}
    _lib.check("resident single-query calls")
    L.destroySuperIntervals(si)
    # python/ctypes call overhead of the same loop shape (a function that returns at once)
    t0 = time.perf_counter()
    for s_, e_ in q:
        L.si_b200_last_error()
    t_ctypes = (time.perf_counter() - t0) / k
    out = {"calls": k, "countOverlaps_us": t_count * 1e6, "searchValues_us": t_search * 1e6, "ctypes_call_overhead_us": t_ctypes * 1e6,
           "hits_per_query": tot / k, "count_equals_search_sizes": tot == tot_v, "resident": res_out,
           "path": "SI_OPT_RESIDENT_QUERIES = 0: query as kernel parameters / mapped pinned mailbox, one launch per call; the kernel publishes a sequence number after its answer and the host spins on that word (csrc/c_abi.cu)"}
    from oracle.pyoracle import Reference
    if Reference.available():
        ref = Reference(starts, ends)
        t1, _ = ref.time_count(qs[:200_000], qe[:200_000], 1)
        out["reference_count_us_one_thread"] = t1 / 200_000 * 1e6
    return out


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

    t_start = time.perf_counter()
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
    is_c2 = a.workload == "c2" and a.intervals == 10_000_000 and a.queries == 100_000_000
    wall = {}

    # ---- synthetic inputs (seeded): same index on every rank, one query shard per rank
    starts, ends, qs, qe = make_workload(a, rank, a.queries)
    nq = qs.size
    first_order = ORDER_SORTED if a.order == "sorted" else ORDER_UNSORTED
    wall["generate"] = time.perf_counter() - t_start

    ix = DeviceIndex()
    if a.algo != "auto":
        from superintervals_b200.device import COUNT_CELLS, COUNT_RANK, COUNT_WALK, OPT_COUNT_ALGO
        ix.set_option(OPT_COUNT_ALGO, {"walk": COUNT_WALK, "rank": COUNT_RANK, "cells": COUNT_CELLS}[a.algo])
    if a.partition:
        ix.set_option(5, 1)
    for env, opt in (("SIB_BUCKET", 1), ("SIB_WSHIFT", 2), ("SIB_GRID", 4), ("SIB_CELLS_FILL", 6)):      # tuning experiments
        if os.environ.get(env):
            ix.set_option(opt, int(os.environ[env]))
    d_s, d_e = torch.from_numpy(starts).cuda(), torch.from_numpy(ends).cuda()
    torch.cuda.synchronize()
    build = bench_build(a, torch, ix, d_s, d_e, rank)
    cells_info = ix.cells_info()
    bits_info = ix.bits_info()
    cells_bytes = cells_info["starts"]["bytes"] + cells_info["ends"]["bytes"]
    d_qs, d_qe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    if a.order == "sorted":
        o = torch.argsort(d_qs, stable=True)
        d_qs, d_qe = d_qs[o].contiguous(), d_qe[o].contiguous()
        qs, qe = d_qs.cpu().numpy(), d_qe.cpu().numpy()
        del o
    counts = torch.empty(nq, dtype=torch.int32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_arm(dq_s, dq_e, out, order, sampler_windows):
        """W warm-up steps, then exactly K steps timed with CUDA events (max over ranks); per-kernel event records."""
        def step():
            ix.count(dq_s, dq_e, out=out, order=order)
        for _ in range(max(a.warmup, 3)):
            step()
        barrier()
        ix.set_option(OPT_TIMING, 1)            # CUDA event pair around every hot kernel launch, on its own stream
        launches0 = L.si_b200_kernel_launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        ev0.record()
        for _ in range(a.steps):
            step()
        ev1.record()
        barrier()
        w1 = time.perf_counter()
        sampler_windows.append((w0, w1))
        ms = ev0.elapsed_time(ev1)
        launches = int(L.si_b200_kernel_launches() - launches0)
        recs = ix.read_timings()
        ix.set_option(OPT_TIMING, 0)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / a.steps, launches, kernel_table(recs, a.steps)

    # ---- timed region: exactly K steps, device timing, clocks sampled meanwhile
    sampler = ClockSampler(local)
    windows = []
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    ms_per_step, launches, kernels = device_arm(d_qs, d_qe, counts, first_order, windows)
    value = world * nq / (ms_per_step * 1e-3)
    passes = int(round(kernels.get("pt_onesweep", {}).get("launches_per_step", 0)))
    dom_tables = (bits_info["bytes"], "its rank bits") if "count_stream" in kernels else (cells_bytes, "the rank cells of both tables")
    roofline = count_roofline({k: v for k, v in kernels.items() if k.startswith("count")}, ms_per_step, nq, dom_tables[0], dom_tables[1], is_c2)

    # ---- per-rank hit totals -> CSR base offsets of each shard (weak scaling: no data-path collective)
    hits = counts.to(torch.int64).sum().reshape(1)
    if world > 1:
        allh = [torch.zeros_like(hits) for _ in range(world)]
        dist.all_gather(allh, hits)
        shard_hits = [int(x.item()) for x in allh]
    else:
        shard_hits = [int(hits.item())]
    shard_base = [int(x) for x in np.concatenate([[0], np.cumsum(shard_hits)[:-1]])]

    # ---- the same batch position-sorted: the streaming kernel (north-star count), device-resident
    srt = None
    s_counts = s_o = None
    if not a.no_sorted and a.order == "shuffled":
        s_o = torch.argsort(d_qs, stable=True)
        sqs, sqe = d_qs[s_o].contiguous(), d_qe[s_o].contiguous()
        s_counts = torch.empty_like(counts)
        s_ms, s_launches, s_kernels = device_arm(sqs, sqe, s_counts, ORDER_SORTED, windows)
        s_tab = (bits_info["bytes"], "its rank bits") if "count_stream" in s_kernels else (cells_bytes, "the rank cells of both tables")
        tiles, handed = ix.stream_stats() if "count_stream" in s_kernels else (0, 0)
        srt = {"workload": "the same batch sorted by query start (stable), device-resident", "value": world * nq / (s_ms * 1e-3), "unit": UNIT,
               "ms_per_step": s_ms, "gpu_launches": s_launches, "kernels": s_kernels,
               "roofline": count_roofline({k: v for k, v in s_kernels.items() if k.startswith("count")}, s_ms, nq, s_tab[0], s_tab[1], is_c2),
               "rank_bits": bits_info, "stream_tiles": tiles, "tiles_handed_to_the_cells_code": handed,
               "equals_shuffled_counts": bool(torch.equal(s_counts, counts[s_o]))}
    wall["device_arms"] = time.perf_counter() - t_start

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (pinned), copies inside the timed region
    si = L.createSuperIntervals()
    L.siSetHostMirror(si, False)                        # queries only: skip the index read-back
    L.addIntervals(si, starts.ctypes.data, ends.ctypes.data, None, starts.size)
    L.indexSuperIntervals(si)
    _lib.check("indexSuperIntervals")
    h_qs, h_qe = torch.from_numpy(qs).pin_memory(), torch.from_numpy(qe).pin_memory()
    h_out = torch.empty(nq, dtype=torch.int64).pin_memory()

    def e2e_arm(fn, steps):
        fn()
        _lib.check("e2e call")
        barrier()
        w2 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        w3 = time.perf_counter()
        windows.append((w2, w3))
        te = torch.tensor([(w3 - w2) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te[0])

    e2e_s = e2e_arm(lambda: L.countOverlapsBatch(si, h_qs.data_ptr(), h_qe.data_ptr(), nq, h_out.data_ptr()), a.e2e_steps)
    e2e_ok = bool((h_out.numpy().astype(np.int64) == counts.cpu().numpy().astype(np.uint32).astype(np.int64)).all())
    ceil64 = copy_ceiling(torch, dist, world, nq, h_qs, h_qe, h_out, 8)
    h_out32 = h_out.view(torch.int32)[:nq]
    e2e32_s = e2e_arm(lambda: L.countOverlapsBatch32(si, h_qs.data_ptr(), h_qe.data_ptr(), nq, h_out32.data_ptr()), a.e2e_steps)
    e2e32_ok = bool(torch.equal(h_out32, counts.cpu()))
    ceil32 = copy_ceiling(torch, dist, world, nq, h_qs, h_qe, h_out, 4)
    e2e = {"value": world * nq / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * nq, "d2h_bytes_per_step": 8 * nq,
           "ms_per_step": e2e_s * 1e3, "pcie_gbs_each_way": 8 * nq / e2e_s / 1e9,
           "call": "countOverlapsBatch(si, qs, qe, n, size_t* counts) with pinned host buffers",
           "copy_ceiling_ms": ceil64 * 1e3, "frac_of_copy_ceiling": ceil64 / e2e_s,
           "copy_ceiling_is": "bare cudaMemcpyAsync of the same pinned buffers, H2D and D2H on two streams at once, every rank at the same time, best of 3",
           "u32_counts": {"value": world * nq / e2e32_s, "unit": UNIT, "ms_per_step": e2e32_s * 1e3, "d2h_bytes_per_step": 4 * nq,
                          "call": "countOverlapsBatch32(si, qs, qe, n, uint32_t* counts)", "equals_device": e2e32_ok,
                          "copy_ceiling_ms": ceil32 * 1e3, "frac_of_copy_ceiling": ceil32 / e2e32_s},
           "rank0_cpus_near_gpu": (len(near_cpus) if near_cpus else None)}
    if srt is not None:
        hs_qs, hs_qe = sqs.cpu().pin_memory(), sqe.cpu().pin_memory()
        s_e2e = e2e_arm(lambda: L.countOverlapsBatch(si, hs_qs.data_ptr(), hs_qe.data_ptr(), nq, h_out.data_ptr()), max(1, a.e2e_steps - 1))
        srt["e2e"] = {"value": world * nq / s_e2e, "unit": UNIT, "ms_per_step": s_e2e * 1e3, "h2d_bytes_per_step": 8 * nq, "d2h_bytes_per_step": 8 * nq,
                      "equals_device": bool((h_out.numpy().astype(np.int64) == s_counts.cpu().numpy().astype(np.uint32).astype(np.int64)).all()),
                      "call": "countOverlapsBatch on the sorted host batch (order detected on the device: SI_ORDER_AUTO)"}
        del hs_qs, hs_qe
    clocks = sampler.stop(windows) if rank == 0 else None
    L.destroySuperIntervals(si)
    del h_out32
    wall["e2e"] = time.perf_counter() - t_start

    def guarded(what, fn):
        """A secondary measurement must not cost the line its headline: a failure is reported in place of the object."""
        try:
            return fn()
        except Exception as ex:   # noqa: BLE001
            import traceback
            sys.stderr.write(f"[bench] {what} failed: {ex!r}\n{traceback.format_exc()}\n")
            try:
                L.si_b200_clear_error()
                torch.cuda.empty_cache()
            except Exception:   # noqa: BLE001
                pass
            return {"error": f"{what}: {ex!r}"[:400]}

    # ---- N > 1: one batch cut N ways, counts all-gathered over NCCL (mode A); search_values with global offsets
    strong = sv = None
    if world > 1:
        rank0_counts = counts.clone()
        dist.broadcast(rank0_counts, 0)
        strong = guarded("strong", lambda: bench_strong(a, torch, dist, L, _lib, ix, world, rank, d_qs, d_qe, rank0_counts))
        del rank0_counts
        if not a.no_search_values:
            sv = guarded("search_values", lambda: bench_search_values_split(a, torch, dist, L, _lib, world, rank))
    elif not a.no_search_values:
        sv = guarded("search_values", lambda: bench_search_values(a, torch, L, _lib, rank))
    wall["search_values"] = time.perf_counter() - t_start

    # keep what the parity checks need, free the big device arrays before the other configs
    counts_head = counts[:2_000_000].cpu().numpy().astype(np.uint32).astype(np.int64)
    s_head = None
    if srt is not None:
        s_head = (sqs[:2_000_000].cpu().numpy(), sqe[:2_000_000].cpu().numpy(), s_counts[:2_000_000].cpu().numpy().astype(np.uint32).astype(np.int64))
        del sqs, sqe, s_counts, s_o
    device_bytes = ix.device_bytes
    del d_qs, d_qe, counts, h_qs, h_qe, h_out
    torch.cuda.empty_cache()

    configs = None
    if not a.no_search_values and not a.no_configs:
        configs = {}
        if world == 1:
            r_ = guarded("configs", lambda: bench_configs(a, torch, dist, L, _lib, world, rank))
            configs.update(r_ if "error" not in r_ else {"c1_c5": r_})
        configs["c4"] = guarded("c4", lambda: bench_c4(a, torch, dist, world, rank))
    wall["configs"] = time.perf_counter() - t_start

    latency = bed = setops = None
    if world == 1 and not a.no_search_values:
        latency = guarded("latency", lambda: bench_latency(a, L, _lib, starts, ends, qs, qe))
        if a.bed_lines > 0:
            bed = guarded("bed_ingest", lambda: bench_bed_ingest(a, rank))
        if a.setop_intervals > 0:
            setops = guarded("set_algebra", lambda: bench_set_algebra(a, L, _lib, rank))
    wall["rows_either_side"] = time.perf_counter() - t_start

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- reference-walk statistics + parity of a sample against the oracle (checker only)
    from oracle.pyoracle import Oracle
    Lg = max(1, math.ceil(math.log2(max(2, a.intervals))))
    stat_n = min(nq, 200_000)
    orc = Oracle(starts, ends)
    h_s, j_s = orc.walk_stats(qs[:stat_n], qe[:stat_n])
    h_per_q = shard_hits[0] / nq                         # exact, from the GPU counts
    j_per_q = j_s / stat_n                               # reference-walk failed tests, sample estimate
    a_walk = 8 + 4 + 4 * Lg + 4 * (h_per_q + j_per_q) + 4 * j_per_q      # SURVEY 8d A_count(q)
    if roofline is not None:
        roofline["whole_step"] = {"compulsory_hbm_bytes": 12 * nq + dom_tables[0], "compulsory_gbs": (12 * nq + dom_tables[0]) / (ms_per_step * 1e-3) / 1e9,
                                  "reference_walk_bytes_per_query": a_walk, "hits_per_query": h_per_q, "jumps_per_query_sampled": j_per_q, "log2_n": Lg,
                                  "note": "SURVEY 8d's A_count describes the reference's element walk; the count kernels answer the closed form "
                                          "#{starts <= qe} - #{ends < qs} and never read those bytes, so no fraction is quoted against it"}
        roofline["note"] = ("kernel_ms: per-launch CUDA events recorded by the library on the launching stream (SI_OPT_TIMING); "
                            "frac = frac_hbm_compulsory; traffic is a committed ncu figure, not measured in this run")
    parity = {"sample": int(stat_n), "mismatches": int((orc.count_batch(qs[:stat_n], qe[:stat_n]).astype(np.int64) != counts_head[:stat_n]).sum()),
              "e2e_equals_device": e2e_ok}
    del orc

    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_reference(a, starts, ends, qs, qe, 3, 1, host_threads())
        sq, se = r["_sample"]
        m = min(sq.size, 2_000_000)
        same = int((r["_counts"](sq[:m], se[:m]).astype(np.int64) != counts_head[:m]).sum())
        parity["reference_sample"] = m
        parity["reference_mismatches"] = same
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_qps", "build_s")}
        if srt is not None:
            # the CPU arm on SORTED input (its fast case), the first cpu_sample queries of the sorted batch
            n_s = min(a.cpu_sample, nq)
            # a sorted sample of the same size: sort the sample itself (a prefix of the sorted batch would span only part of the axis)
            r2 = cpu_reference(a, starts, ends, qs, qe, 2, 1, host_threads(), order="sorted", ref=r["_ref"])
            srt["cpu_baseline"] = {k: r2[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_qps")}
            mm = min(s_head[0].size, 2_000_000)
            srt["parity"] = {"reference_sample": mm, "reference_mismatches": int((r["_counts"](s_head[0][:mm], s_head[1][:mm]).astype(np.int64) != s_head[2][:mm]).sum())}
    wall["cpu_baseline"] = time.perf_counter() - t_start

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(a), "intervals": a.intervals, "queries_per_gpu": nq,
                       "step": ((f"device partition of the batch by position ({passes} onesweep passes over 12-byte records) + count kernel"
                                 if passes else "count kernel straight on the shuffled batch (rank cells resident in L2, no partition)")
                                if a.order == "shuffled" else "count kernel on position-sorted queries"),
                       "count_kernel": next((k for k in ("count_cells", "count_stream", "count_rank", "count_walk") if k in kernels), None),
                       "rank_cells": cells_info,
                       "count_algo": a.algo,
                       "l2": (f"inputs larger than L2: {8 * nq / 1e6:.0f} MB of queries + {4 * nq / 1e6:.0f} MB of counts per step vs 126 MB"
                              if 12 * nq > 2 * 126e6 else
                              f"inputs ({12 * nq / 1e6:.0f} MB per step) FIT in L2: a parity/side configuration, not the metric's workload"),
                       "index": "replicated per GPU",
                       "collective": ("none in the weak-scaling headline (independent shards); the `strong` object cuts ONE batch N ways and all-gathers the "
                                      "per-query counts over NCCL" if world > 1 else "none at N = 1")},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "sorted": srt, "strong": strong,
            "cpu_baseline": cpu_baseline, "parity": parity, "build": build, "build_ms": build["ms"],
            "shard_hits": shard_hits, "shard_csr_base": shard_base, "device_bytes": device_bytes,
            "search_values": sv, "configs": configs, "latency": latency, "bed_ingest": bed, "set_algebra": setops,
            "wall_s": {k: round(v, 2) for k, v in wall.items()}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
