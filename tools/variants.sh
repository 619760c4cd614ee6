# Times the count bench with alternative builds of the library (superintervals_b200/variants/lib_*.so,
# made with `make -C superintervals_b200/csrc variant NAME=.. DEFS=..`). usage: bash tools/variants.sh [bench args]
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d" % d["parity"]["mismatches"])'
echo "== default"; python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "$show"
for f in superintervals_b200/variants/lib_*.so; do
  echo "== $f"; SIB_LIBRARY=$PWD/$f python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "$show"
done
