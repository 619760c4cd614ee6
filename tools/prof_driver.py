"""Tiny driver for ncu captures: builds one index, runs the hot kernels a few times.
usage: prof_driver.py [c1|c2|c2s|c3] [count|count_unsorted|count_walk|search] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superintervals_b200 import workloads as W
from superintervals_b200.device import (DeviceIndex, ORDER_SORTED, ORDER_UNSORTED, OPT_COUNT_ALGO, COUNT_WALK,
                                        OPT_BUCKET_INTERVALS, OPT_WINDOW_SHIFT)

if len(sys.argv) > 1 and sys.argv[1] == "mixed":
    # mode B in one launch: 8 contigs of 2 M read-length intervals on 50 Mb each, 64 M mixed queries
    import numpy as np
    from superintervals_b200.genome import GenomeIndex
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    gi = GenomeIndex([f"c{i}" for i in range(8)], [2_000_000] * 8, rank=0, world=1)
    for c in range(8):
        s, e = W.config2_intervals(2_000_000, 20 + c, axis=50_000_000)
        gi.build_contig(c, torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
    qs, qe = W.config2_queries(64_000_000, 3, axis=50_000_000)
    cid = torch.from_numpy(np.random.default_rng(1).integers(0, 8, qs.size).astype(np.int32)).cuda()
    dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    for _ in range(reps):
        out = gi.count_mixed(cid, dqs, dqe)
    torch.cuda.synchronize()
    print("hits", int(out.long().sum().item()))
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "c5lite":
    # C5's density at 1/32 of its size: 32 M intervals on a 64 Mb axis, 64 M stabbing queries (the pair cells answer them)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    n5 = 32_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    st = (torch.rand(n5, generator=g, device="cuda", dtype=torch.float64) * 2 * n5).to(torch.int64)
    ln = (150 + torch.rand(n5, generator=g, device="cuda", dtype=torch.float64) * (10_000 - 150)).to(torch.int64)
    ix = DeviceIndex().build(st.to(torch.int32), torch.clamp(st + ln, max=2**31 - 1).to(torch.int32))
    q = (torch.rand(2 * n5, generator=g, device="cuda", dtype=torch.float64) * 2 * n5).to(torch.int64).to(torch.int32)
    out = torch.empty_like(q)
    for _ in range(reps):
        ix.count(q, q, out=out, order=ORDER_UNSORTED)
    torch.cuda.synchronize()
    print("hits", int(out.long().sum().item()), ix.cells_info()["pair"])
    sys.exit(0)
which = sys.argv[1] if len(sys.argv) > 1 else "c2s"
mode = sys.argv[2] if len(sys.argv) > 2 else "count"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
gen = {"c1": lambda: W.config1(1_000_000, 0),
       "c2s": lambda: W.config2(1_000_000, 10_000_000, 2, axis=25_000_000),
       "c2": lambda: W.config2(10_000_000, 100_000_000, 2),
       "c3": lambda: W.config3()}[which]
s, e, qs, qe = gen()
ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
if os.environ.get("SIB_BUCKET"):
    ix.set_option(OPT_BUCKET_INTERVALS, int(os.environ["SIB_BUCKET"]))
if os.environ.get("SIB_WSHIFT"):
    ix.set_option(OPT_WINDOW_SHIFT, int(os.environ["SIB_WSHIFT"]))
dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
out = torch.empty_like(dqs)
if mode in ("count", "count_walk_sorted", "search_sorted"):
    order = torch.argsort(dqs, stable=True)
    sqs, sqe = dqs[order].contiguous(), dqe[order].contiguous()
if mode.startswith("count_walk"):
    ix.set_option(OPT_COUNT_ALGO, COUNT_WALK)
for _ in range(reps):
    if mode in ("count", "count_walk_sorted"):
        ix.count(sqs, sqe, out=out, order=ORDER_SORTED)
    elif mode in ("count_unsorted", "count_walk"):
        ix.count(dqs, dqe, out=out, order=ORDER_UNSORTED)
    elif mode == "search_sorted":
        ix.search_values(sqs, sqe, order=ORDER_SORTED)
    else:
        ix.search_values(dqs, dqe, order=ORDER_UNSORTED)
torch.cuda.synchronize()
print("hits", int(out.long().sum().item()))
