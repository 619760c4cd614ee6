"""Per-SASS-instruction stall samples of one kernel in an .ncu-rep (read here, no GPU needed):
prints the top instructions by sampled stalls, with the dominant stall reasons.
usage: ncu_source.py report.ncu-rep [kernel-name-regex] [top] [launch-skip]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else "."; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kid}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[idx["# Samples"]])
    except ValueError: continue
    tot += n
    st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    data.append((n, r[idx["Address"]][-5:], r[idx["Source"]][:90], st, r[idx["Instructions Executed"]]))
print("total samples", tot)
cum = 0
for i, (n, addr, src, st, ex) in enumerate(data):
    data[i] = (n, addr, src, st, ex, i)
for n, addr, src, st, ex, i in sorted(data, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}%  #{i:4d} {src:90s} exec={ex:>9s} " + " ".join(f"{k}:{v}" for v, k in st if v))
