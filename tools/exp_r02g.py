"""r02g experiments: (1) what the persisting-L2 window buys the cells count and costs the fill that follows;
(2) C5-lite: rank cells larger than L2 gathered straight from HBM vs behind the locality partition.
usage: exp_r02g.py [persist|c5|build]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superintervals_b200 import workloads as W
from superintervals_b200.device import (DeviceIndex, ORDER_SORTED, ORDER_UNSORTED, OPT_TIMING, OPT_CELLS_DIRECT_BYTES)


def timed(fn, steps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def kernels(ix, steps):
    out = {}
    for tag, ms in ix.read_timings():
        out.setdefault(tag, []).append(ms)
    return {k: round(sum(v) / steps, 4) for k, v in out.items()}


what = sys.argv[1] if len(sys.argv) > 1 else "persist"
res = {"what": what, "SIB_L2_PERSIST": os.environ.get("SIB_L2_PERSIST")}
if what == "persist":
    s3, e3, qs3, qe3 = W.config3()
    ix3 = DeviceIndex().build(torch.from_numpy(s3).cuda(), torch.from_numpy(e3).cuda())
    d3s, d3e = torch.from_numpy(qs3).cuda(), torch.from_numpy(qe3).cuda()
    off, vals = ix3.search_values(d3s, d3e, order=ORDER_UNSORTED)
    cnt3 = torch.empty_like(d3s)

    def sv():
        ix3.count(d3s, d3e, out=cnt3, order=ORDER_UNSORTED)
        ix3.search_values(d3s, d3e, order=ORDER_UNSORTED, counts=cnt3, offsets=off, out=vals)
    for _ in range(3):
        sv()
    ix3.set_option(OPT_TIMING, 1)
    res["c3_fresh_ms"] = timed(sv, 10)
    res["c3_fresh_kernels"] = kernels(ix3, 10)
    ix3.set_option(OPT_TIMING, 0)
    s, e, qs, qe = W.config2(10_000_000, 100_000_000, 2)
    ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    out = torch.empty_like(dqs)
    for _ in range(3):
        ix.count(dqs, dqe, out=out, order=ORDER_UNSORTED)
    res["c2_shuffled_ms"] = timed(lambda: ix.count(dqs, dqe, out=out, order=ORDER_UNSORTED), 10)
    o = torch.argsort(dqs, stable=True)
    sqs, sqe = dqs[o].contiguous(), dqe[o].contiguous()
    for _ in range(3):
        ix.count(sqs, sqe, out=out, order=ORDER_SORTED)
    res["c2_sorted_ms"] = timed(lambda: ix.count(sqs, sqe, out=out, order=ORDER_SORTED), 10)
    ix.count(dqs, dqe, out=out, order=ORDER_UNSORTED)          # leaves C2's cells in L2 as the bench does
    for _ in range(3):
        sv()
    ix3.set_option(OPT_TIMING, 1)
    res["c3_after_c2_ms"] = timed(sv, 10)
    res["c3_after_c2_kernels"] = kernels(ix3, 10)
elif what == "c5":
    n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 32_000_000
    axis = int(2_000_000_000 * (n5 / 1_000_000_000))
    g = torch.Generator(device="cuda").manual_seed(5)
    st = (torch.rand(n5, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64)
    ln = (150 + torch.rand(n5, generator=g, device="cuda", dtype=torch.float64) * (10_000 - 150)).to(torch.int64)
    s5, e5 = st.to(torch.int32), torch.clamp(st + ln, max=2**31 - 1).to(torch.int32)
    del st, ln
    ix = DeviceIndex().build(s5, e5)
    res["cells"] = ix.cells_info()
    q5 = (torch.rand(2 * n5, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64).to(torch.int32)
    c_part = torch.empty(2 * n5, dtype=torch.int32, device="cuda")
    c_dir = torch.empty_like(c_part)
    for _ in range(2):
        ix.count(q5, q5, out=c_part, order=ORDER_UNSORTED)
    ix.set_option(OPT_TIMING, 1)
    res["partition_ms"] = timed(lambda: ix.count(q5, q5, out=c_part, order=ORDER_UNSORTED), 5)
    res["partition_kernels"] = kernels(ix, 5)
    ix.set_option(OPT_CELLS_DIRECT_BYTES, 1 << 40)
    for _ in range(2):
        ix.count(q5, q5, out=c_dir, order=ORDER_UNSORTED)
    res["direct_ms"] = timed(lambda: ix.count(q5, q5, out=c_dir, order=ORDER_UNSORTED), 5)
    res["direct_kernels"] = kernels(ix, 5)
    res["equal"] = bool(torch.equal(c_part, c_dir))
    sq = torch.sort(q5).values
    for _ in range(2):
        ix.count(sq, sq, out=c_dir, order=ORDER_SORTED)
    res["sorted_ms"] = timed(lambda: ix.count(sq, sq, out=c_dir, order=ORDER_SORTED), 5)
    res["sorted_kernels"] = kernels(ix, 5)
    res["bits"] = ix.bits_info()
elif what == "fill":
    # C3 search_values with the (value, end) lists on and off: per-kernel times and equality of the results
    from superintervals_b200._lib import OPT_STAB_VALUE_LISTS
    s3, e3, qs3, qe3 = W.config3()
    d3s, d3e = torch.from_numpy(qs3).cuda(), torch.from_numpy(qe3).cuda()
    keep = {}
    for vl in (1, 0):
        ix3 = DeviceIndex()
        ix3.set_option(OPT_STAB_VALUE_LISTS, vl)
        ix3.build(torch.from_numpy(s3).cuda(), torch.from_numpy(e3).cuda())
        off, vals = ix3.search_values(d3s, d3e, order=ORDER_UNSORTED)
        cnt3 = torch.empty_like(d3s)

        def sv():
            ix3.count(d3s, d3e, out=cnt3, order=ORDER_UNSORTED)
            ix3.search_values(d3s, d3e, order=ORDER_UNSORTED, counts=cnt3, offsets=off, out=vals)
        for _ in range(3):
            sv()
        ix3.set_option(OPT_TIMING, 1)
        res[f"vlists{vl}_ms"] = timed(sv, 10)
        res[f"vlists{vl}_kernels"] = kernels(ix3, 10)
        res[f"vlists{vl}_bytes"] = ix3.device_bytes
        keep[vl] = (off.clone(), vals.clone())
        del ix3
    res["equal"] = bool(torch.equal(keep[0][0], keep[1][0]) and torch.equal(keep[0][1], keep[1][1]))
elif what == "mixed":
    import numpy as np
    from superintervals_b200.genome import GenomeIndex
    gi = GenomeIndex([f"c{i}" for i in range(8)], [2_000_000] * 8, rank=0, world=1)
    for c in range(8):
        s, e = W.config2_intervals(2_000_000, 20 + c, axis=50_000_000)
        gi.build_contig(c, torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    qs, qe = W.config2_queries(64_000_000, 3, axis=50_000_000)
    cid = torch.from_numpy(np.random.default_rng(1).integers(0, 8, qs.size).astype(np.int32)).cuda()
    dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    for _ in range(3):
        out = gi.count_mixed(cid, dqs, dqe)
    res["mixed_ms"] = timed(lambda: gi.count_mixed(cid, dqs, dqe), 10)
    res["hits"] = int(out.long().sum().item())
    res["lib"] = os.environ.get("SIB_LIBRARY")
else:
    s, e = W.config2_intervals(10_000_000, 2)
    ds, de = torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda()
    ix = DeviceIndex()
    for _ in range(3):
        ix.build(ds, de)
    res["build_ms"] = timed(lambda: ix.build(ds, de), 3)
print(json.dumps(res))
