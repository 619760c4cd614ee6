# sweep partition tunables on the bench workload. usage: bash tools/sweep.sh
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f x%g" % (n, v["ms_per_launch"], v["launches_per_step"]) for n, v in k.items()) + " | mismatches %d" % d["parity"]["mismatches"])'
for w in 22 23 24 25; do for b in 1024 256; do
  echo "== wshift $w bucket $b"; SIB_WSHIFT=$w SIB_BUCKET=$b python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "$show"
done; done
