"""Tiny driver for ncu captures: builds one index, runs the hot kernels a few times.
usage: prof_driver.py [c1|c2|c2s|c3] [count|search] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superintervals_b200 import workloads as W
from superintervals_b200.device import DeviceIndex, ORDER_SORTED, ORDER_UNSORTED

which = sys.argv[1] if len(sys.argv) > 1 else "c2s"
mode = sys.argv[2] if len(sys.argv) > 2 else "count"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
gen = {"c1": lambda: W.config1(1_000_000, 0),
       "c2s": lambda: W.config2(1_000_000, 10_000_000, 2, axis=25_000_000),
       "c2": lambda: W.config2(10_000_000, 100_000_000, 2),
       "c3": lambda: W.config3()}[which]
s, e, qs, qe = gen()
ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
order = torch.argsort(dqs, stable=True)
sqs, sqe = dqs[order].contiguous(), dqe[order].contiguous()
out = torch.empty_like(sqs)
for _ in range(reps):
    if mode == "count":
        ix.count(sqs, sqe, out=out, order=ORDER_SORTED)
    elif mode == "count_unsorted":
        ix.count(dqs, dqe, out=out, order=ORDER_UNSORTED)
    else:
        ix.search_values(sqs, sqe, order=ORDER_SORTED)
torch.cuda.synchronize()
print("hits", int(out.long().sum().item()))
