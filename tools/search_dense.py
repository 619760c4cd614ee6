"""search_idxs (CSR) on a C2-dense index at scale: 10 M read-length intervals x Q range queries on the
250 Mb axis (~137 hits per query, ~96 of them intervals that start before the query: the stab part).
Too many hits for the oracle, so the result is checked through size-independent properties:
counts equal the closed form, every list is strictly descending in position, every listed
interval overlaps its query. One JSON line.   usage: python tools/search_dense.py [queries]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from superintervals_b200 import workloads as W
from superintervals_b200.device import DeviceIndex, FILL_IDXS, OPT_TIMING, ORDER_UNSORTED

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
s, e = W.config2_intervals(10_000_000, 2)
qs, qe = W.config2_queries(nq, 2, shard=0)
ix = DeviceIndex().build(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda())
dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
off, idx = ix.search(dqs, dqe, FILL_IDXS, order=ORDER_UNSORTED)
total = int(off[nq].item())
counts = torch.empty(nq, dtype=torch.int32, device="cuda")
ix.set_option(OPT_TIMING, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record()
for _ in range(steps):
    ix.search(dqs, dqe, FILL_IDXS, order=ORDER_UNSORTED, counts=counts, offsets=off, out=idx)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
per = {}
for name, t in ix.read_timings():
    per[name] = per.get(name, 0.0) + t / steps
# properties
v = ix.view() if hasattr(ix, "view") else None
ds, de = torch.from_numpy(np.sort(s)).cuda(), torch.from_numpy(np.sort(e)).cuda()
closed = torch.searchsorted(ds, dqe, right=True) - torch.searchsorted(de, dqs, right=False)
cnt = (off[1:] - off[:-1])
ok_counts = bool((cnt == closed).all())
srt = ix.sorted_arrays() if hasattr(ix, "sorted_arrays") else None
qid = torch.repeat_interleave(torch.arange(nq, device="cuda"), cnt)
idx64 = idx.to(torch.int64)
desc = bool(((idx64[1:] < idx64[:-1]) | (qid[1:] != qid[:-1])).all())
# position order = (start asc, end desc, insertion): rebuild it with a stable sort to test the overlaps
order = np.lexsort((np.arange(s.size), -e.astype(np.int64), s))
ps, pe = torch.from_numpy(s[order]).cuda(), torch.from_numpy(e[order]).cuda()
overlap = bool(((ps[idx64] <= dqe[qid]) & (pe[idx64] >= dqs[qid])).all())
print(json.dumps({"workload": f"C2-dense search_idxs: 10M intervals x {nq/1e6:g}M shuffled queries", "hits": total,
                  "hits_per_query": total / nq, "ms_per_step": ms, "queries_per_s": nq / (ms * 1e-3),
                  "hits_per_s": total / (ms * 1e-3), "kernel_ms_per_step": per, "stab_lists": ix.stab_info(),
                  "result_gbs": (4 * total + 8 * nq) / (ms * 1e-3) / 1e9,
                  "properties": {"counts_equal_closed_form": ok_counts, "lists_strictly_descending": desc,
                                 "every_listed_interval_overlaps": overlap}}))
