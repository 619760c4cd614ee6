# count kernel iteration: parity + bench of the count path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py -m gpu -x -q > gpurun_out/tests_count.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_count.log; tail -8 gpurun_out/tests_count.log
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d ref %s" % (d["parity"]["mismatches"], d["parity"].get("reference_mismatches")), d["config"].get("rank_cells"))'
echo "== shuffled"; timeout 300 python bench.py --no-search-values --e2e-steps 1 --steps 5 2>&1 | tail -1 | python -c "$show"
echo "== sorted"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 --steps 5 --order sorted 2>&1 | tail -1 | python -c "$show"
