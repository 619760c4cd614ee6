# round 2, first GPU session: streaming kernel parity + timing, L2 gather ceiling, count-cells variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q > gpurun_out/r02a_stream_tests.log 2>&1; echo "stream tests rc=$?" 
tail -5 gpurun_out/r02a_stream_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py -x -q > gpurun_out/r02a_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/r02a_parity.log
timeout 300 tools/bin/l2_peak > gpurun_out/l2_peak.json 2> gpurun_out/l2_peak.err; echo "l2_peak rc=$?"
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d e2e %.2f ms" % (d["parity"]["mismatches"], d["e2e"]["ms_per_step"]))'
echo "== sorted (stream)"; timeout 300 python bench.py --order sorted --no-cpu-baseline --no-search-values --e2e-steps 2 --steps 10 2> gpurun_out/r02a_sorted.err | tail -1 | tee gpurun_out/r02a_sorted.json | python -c "$show"
echo "== shuffled (base)"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 2 --steps 10 2> gpurun_out/r02a_shuf.err | tail -1 | tee gpurun_out/r02a_shuf.json | python -c "$show"
for f in superintervals_b200/variants/lib_*.so; do
  echo "== $f"; SIB_LIBRARY=$PWD/$f timeout 200 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 --steps 5 2>&1 | tail -1 | python -c "$show"
done
