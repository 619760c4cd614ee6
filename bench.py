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

value   queries/s, whole job, queries resident in HBM (shuffled order: the step
        includes the device radix sort of the batch by start + the count kernel).
e2e     same metric through the C-ABI host-buffer call countOverlapsBatch()
        (pinned host buffers; H2D of the queries and D2H of the counts inside).
roofline / cpu_baseline: see DESIGN.md section "Measurement".
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
    ap.add_argument("--cpu-sample", type=int, default=16_000_000, help="queries timed on the host CPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"C2: {a.intervals/1e6:g}M read-length intervals (150bp-10kb) x {a.queries/1e6:g}M range queries "
            f"per GPU, 250Mb axis, count only, {a.order} queries")


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
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


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
    from superintervals_b200 import workloads as W
    starts, ends = W.config2_intervals(a.intervals, 2)
    qs, qe = W.config2_queries(min(a.queries, a.cpu_sample), 2, shard=0)
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
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist
    from superintervals_b200 import _lib, workloads as W
    from superintervals_b200.device import DeviceIndex, ORDER_SORTED, ORDER_UNSORTED

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: superintervals_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()

    # ---- synthetic inputs (seeded): same index on every rank, one query shard per rank
    starts, ends = W.config2_intervals(a.intervals, 2)
    qs, qe = W.config2_queries(a.queries, 2, shard=rank)
    if a.order == "sorted":
        o = np.argsort(qs, kind="stable")
        qs, qe = np.ascontiguousarray(qs[o]), np.ascontiguousarray(qe[o])
    nq = qs.size
    order = ORDER_SORTED if a.order == "sorted" else ORDER_UNSORTED

    ix = DeviceIndex()
    d_s, d_e = torch.from_numpy(starts).cuda(), torch.from_numpy(ends).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ix.build(d_s, d_e); torch.cuda.synchronize(); build_ms = (time.perf_counter() - t0) * 1e3
    d_qs, d_qe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    counts = torch.empty(nq, dtype=torch.int32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(evs=None):
        if order == ORDER_UNSORTED:
            ix.sort_queries(d_qs)                    # radix sort of the batch by start (arms the next count)
        if evs is not None:
            evs[0].record()
        ix.count(d_qs, d_qe, out=counts, order=order)
        if evs is not None:
            evs[1].record()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, device timing, clocks sampled meanwhile
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = L.si_b200_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    ev0.record()
    for k in range(a.steps):
        step(kev[k])
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = int(L.si_b200_kernel_launches() - launches0)
    count_kernel_ms = float(np.mean([x.elapsed_time(y) for x, y in kev]))
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([elapsed_ms, count_kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, count_kernel_ms_max = float(t[0]), float(t[1])
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
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / a.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e_ok = bool((h_out.numpy().astype(np.int64) == counts.cpu().numpy().astype(np.uint32).astype(np.int64)).all())
    L.destroySuperIntervals(si)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (qk_count_kernel): algorithmic bytes / its own duration
    from oracle.pyoracle import Oracle   # checker only: walk statistics + parity of a sample
    Lg = max(1, math.ceil(math.log2(max(2, a.intervals))))
    stat_n = min(nq, 200_000)
    orc = Oracle(starts, ends)
    h_s, j_s = orc.walk_stats(qs[:stat_n], qe[:stat_n])
    h_per_q = shard_hits[0] / nq                         # exact, from the GPU counts
    j_per_q = j_s / stat_n                               # reference-walk failed tests, sample estimate
    a_count = 8 + 4 + 4 * Lg + 4 * (h_per_q + j_per_q) + 4 * j_per_q      # SURVEY 8d A_count(q)
    peak, peak_src = measured_peak()
    achieved = nq * a_count / (count_kernel_ms_max * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "count_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "qk_count_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                "kernel_ms": count_kernel_ms_max, "algorithmic_bytes_per_query": a_count,
                "hits_per_query": h_per_q, "jumps_per_query_sampled": j_per_q, "log2_n": Lg,
                "compulsory_hbm_bytes_per_launch": 12 * nq + 12 * a.intervals,
                "note": "algorithmic bytes are the reference walk's element-granular traffic (no cache credit); "
                        "the kernel serves them from L1/L2, so frac can exceed 1 -- see traffic for DRAM bytes"}

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
                       "step": "device radix sort of the batch by start + count kernel" if a.order == "shuffled"
                               else "count kernel on position-sorted queries",
                       "l2": "inputs larger than L2: 800 MB of queries + 400 MB of counts per step vs 126 MB",
                       "index": "replicated per GPU", "collective": "all_gather of per-rank hit totals (CSR bases)"},
            "e2e": {"value": world * nq / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * nq,
                    "d2h_bytes_per_step": 8 * nq, "ms_per_step": e2e_s * 1e3,
                    "call": "countOverlapsBatch(si, qs, qe, n, size_t* counts) with pinned host buffers"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity": parity, "build_ms": build_ms, "count_kernel_ms": count_kernel_ms_max,
            "shard_hits": shard_hits, "shard_csr_base": shard_base,
            "device_bytes": ix.device_bytes}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
