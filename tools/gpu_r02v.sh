# round 2, twenty-second GPU session: value lists for the CSR fill, CTA shapes of the rank-cells kernel, suite
mkdir -p gpurun_out
echo "== fill value lists"; timeout 300 python tools/exp_r02g.py fill | tail -1 | tee gpurun_out/r02v_fill.json
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.4f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.4f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d" % (d["parity"]["mismatches"]))'
echo "== default"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 1 --steps 20 2>/dev/null | tail -1 | python -c "$show"
for f in superintervals_b200/variants/lib_qc_*.so; do
  echo "== $f"; SIB_LIBRARY=$PWD/$f timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 1 --steps 20 2>/dev/null | tail -1 | python -c "$show"
done
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02v_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02v_tests.log
