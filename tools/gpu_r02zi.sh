# round 2, session zi, zj: streaming count kernel CTA shapes (C2 sorted)
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.4f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.4f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d ref %s" % (d["parity"]["mismatches"], d["parity"].get("reference_mismatches")))'
echo "== default"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 1 --steps 20 --order sorted 2>/dev/null | tail -n 1 | python -c "$show"
for f in superintervals_b200/variants/lib_sk*.so; do
  echo "== $f"; SIB_LIBRARY=$PWD/$f timeout 300 python bench.py --no-cpu-baseline --no-search-values --no-sorted --e2e-steps 1 --steps 20 --order sorted 2>/dev/null | tail -n 1 | python -c "$show"
done
( timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q ) 2>&1 | tail -n 3
for f in superintervals_b200/variants/lib_sk*.so; do echo "== tests $f"; ( SIB_LIBRARY=$PWD/$f timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q ) 2>&1 | tail -n 2; done
