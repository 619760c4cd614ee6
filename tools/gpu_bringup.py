"""First-light diagnostics on a real GPU: parity against the oracle with mismatch
details, then rough timings. Scratch tool (bench.py is the measured path)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.pyoracle import Oracle
from superintervals_b200 import workloads as W, IntervalMap, _lib
from superintervals_b200.device import DeviceIndex, ORDER_SORTED, ORDER_UNSORTED, ORDER_ASIS

print(torch.cuda.get_device_name(0), "lib", _lib.lib().si_b200_version())

def diff(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}"); return False
    bad = np.flatnonzero(a != b) if a.ndim == 1 else np.flatnonzero((a != b).any(axis=1))
    if bad.size:
        print(f"  {name}: {bad.size} mismatches, first at {bad[:5]}: got {a[bad[:5]]} want {b[bad[:5]]}")
        return False
    print(f"  {name}: ok ({a.shape[0]})")
    return True

ok = True
for label, (s, e, qs, qe) in {
    "c1-20k": W.config1(20_000, 0),
    "c1-1M": W.config1(1_000_000, 0),
    "c3-200k": W.config3(200_000, 200_000, 42, axis=12_000_000),
    "c2-small": W.config2(100_000, 1_000_000, 2, axis=2_500_000),
}.items():
    print(label)
    o = Oracle(s, e)
    t = time.time(); m = IntervalMap.from_arrays(s, e); print(f"  build {1e3*(time.time()-t):.1f} ms")
    ok &= diff("starts", m.starts, o.starts)
    ok &= diff("ends", m.ends, o.ends)
    ok &= diff("data", m.data_index, o.data)
    ok &= diff("branch", m.branch, o.branch)
    want = o.count_batch(qs, qe)
    t = time.time(); got = m.count_batch_np(qs, qe); print(f"  count_batch(host) {1e3*(time.time()-t):.1f} ms")
    ok &= diff("count", got, want)
    off_o, res = o.search_batch(qs, qe, want=("values", "idxs", "keys"))
    t = time.time(); off, vals = m.search_values_batch_csr(qs, qe); print(f"  search_values(host) {1e3*(time.time()-t):.1f} ms, hits {len(vals)}")
    ok &= diff("offsets", off, off_o)
    ok &= diff("values", vals, res["values"])
    _, keys = m.search_keys_batch_csr(qs, qe)
    ok &= diff("keys", keys, res["keys"])
    ok &= diff("any", m.has_overlaps_batch(qs, qe), o.has_overlaps_batch(qs, qe))
print("PARITY", "OK" if ok else "FAILED")

# ---- rough device-resident timings ------------------------------------------------------
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts = []
    for _ in range(reps):
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    return min(ts), float(np.median(ts))

for label, gen in {"C1 1M x 1M": lambda: W.config1(1_000_000, 0),
                   "C2/10 1M x 10M": lambda: W.config2(1_000_000, 10_000_000, 2, axis=25_000_000),
                   "C2 10M x 100M": lambda: W.config2(10_000_000, 100_000_000, 2),
                   "C3 4M x 4M": lambda: W.config3()}.items():
    s, e, qs, qe = gen()
    ds, de = torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda()
    ix = DeviceIndex()
    tb = timeit(lambda: ix.build(ds, de), reps=3)
    dqs, dqe = torch.from_numpy(qs).cuda(), torch.from_numpy(qe).cuda()
    order = torch.argsort(dqs, stable=True)
    sqs, sqe = dqs[order].contiguous(), dqe[order].contiguous()
    out = torch.empty_like(dqs)
    t_sorted = timeit(lambda: ix.count(sqs, sqe, out=out, order=ORDER_SORTED))
    tot_sorted = int(out.long().sum().item())
    t_unsorted = timeit(lambda: ix.count(dqs, dqe, out=out, order=ORDER_UNSORTED))
    tot_unsorted = int(out.long().sum().item())
    t_asis = timeit(lambda: ix.count(dqs, dqe, out=out, order=ORDER_ASIS))
    nq = qs.size
    print(f"{label}: build {tb[0]:.2f} ms | count sorted {t_sorted[0]:.3f} ms ({nq/t_sorted[0]/1e6:.2f} Gq/s) "
          f"| unsorted(+sort) {t_unsorted[0]:.3f} ms ({nq/t_unsorted[0]/1e6:.2f} Gq/s) | as-is {t_asis[0]:.3f} ms "
          f"| hits {tot_sorted} {tot_unsorted}")
    if nq <= 10_000_000:
        offs, vals = ix.search_values(sqs, sqe, order=ORDER_SORTED)
        torch.cuda.synchronize()
        cnt = ix.count(sqs, sqe, order=ORDER_SORTED)
        offs = ix.scan(cnt)
        vals = torch.empty(int(offs[-1].item()), dtype=torch.int32, device="cuda")
        t_scan = timeit(lambda: ix.scan(cnt, out=offs))
        L = _lib.lib()
        import ctypes as C
        def fill():
            L.siFillDevice(ix._ix, sqs.data_ptr(), sqe.data_ptr(), nq, offs.data_ptr(), 0, vals.data_ptr(), ORDER_SORTED,
                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
        t_fill = timeit(fill)
        print(f"    scan {t_scan[0]:.3f} ms | fill {t_fill[0]:.3f} ms | hits {vals.numel()} -> search_values "
              f"{nq/(t_sorted[0]+t_scan[0]+t_fill[0])/1e6:.2f} Gq/s")
    del ix
print("launches", _lib.lib().si_b200_kernel_launches())
