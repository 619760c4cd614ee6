# round 2, twenty-first GPU session: variants of the rank-cells count kernel on C2 shuffled (L2 eviction hints, CTA shapes)
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.4f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.4f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d" % (d["parity"]["mismatches"]))'
echo "== default"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 1 --steps 20 2>/dev/null | tail -1 | python -c "$show"
for f in superintervals_b200/variants/lib_qc_*.so; do
  echo "== $f"; SIB_LIBRARY=$PWD/$f timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 1 --steps 20 2>/dev/null | tail -1 | python -c "$show"
done
( time timeout 600 python -m pytest tests/test_genome.py -m gpu -q -x ) 2>&1 | tail -3
