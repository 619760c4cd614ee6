"""build() probe: C5-density intervals generated on the device (n intervals, 150 bp-10 kb, on an axis of 2n bp,
capped below 2^31), build timed with CUDA events (best of reps), per-kernel times from the library's own events.
usage: build_probe.py n [reps] [c2]     (c2: C2's law instead -- 250 Mb axis scaled with n / 10 M)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superintervals_b200.device import DeviceIndex, OPT_TIMING

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
law = sys.argv[3] if len(sys.argv) > 3 else "c5"
axis = min(2 * n, 2_000_000_000) if law == "c5" else min(25 * n, 2_000_000_000)
g = torch.Generator(device="cuda").manual_seed(5)
st = (torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * axis).to(torch.int64)
ln = (150 + torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * (10_000 - 150)).to(torch.int64)
s, e = st.to(torch.int32), torch.clamp(st + ln, max=2**31 - 1).to(torch.int32)
del st, ln
ix = DeviceIndex()
ix.build(s, e)
torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ix.build(s, e); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
res = {"n": n, "axis": axis, "law": law, "build_ms": best, "intervals_per_s": n / (best * 1e-3), "lib": os.environ.get("SIB_LIBRARY"),
       "device_bytes": ix.device_bytes}
ix.set_option(OPT_TIMING, 1)
ix.build(s, e); torch.cuda.synchronize()
k = {}
for tag, ms in ix.read_timings():
    k.setdefault(tag, [0, 0.0]); k[tag][0] += 1; k[tag][1] += ms
res["kernels"] = {t: {"launches": c, "ms": round(m, 4)} for t, (c, m) in k.items()}
print(json.dumps(res))
