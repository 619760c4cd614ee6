# times alternative builds of the count kernel (superintervals_b200/variants/lib_*.so), each under a short timeout
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d" % d["parity"]["mismatches"])'
for f in superintervals_b200/variants/lib_*.so; do
  echo "== $f"; SIB_LIBRARY=$PWD/$f timeout 150 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 --steps 5 2>&1 | tail -1 | python -c "$show"
done > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
