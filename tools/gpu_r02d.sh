# round 2, fourth GPU session: persistent pipelined streaming kernel; multi-device entry points on one device
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_multi.py tests/test_gpu_device.py -q -x > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02d_tests.log
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d e2e %.2f ms" % (d["parity"]["mismatches"], d["e2e"]["ms_per_step"]))'
echo "== sorted (stream, default)"; timeout 300 python bench.py --order sorted --no-cpu-baseline --no-search-values --e2e-steps 2 --steps 10 2> gpurun_out/r02d_sorted.err | tail -1 | tee gpurun_out/r02d_sorted.json | python -c "$show"
for f in superintervals_b200/variants/lib_sk*.so; do
  echo "== $f sorted"; SIB_LIBRARY=$PWD/$f timeout 200 python bench.py --order sorted --no-cpu-baseline --no-search-values --e2e-steps 1 --steps 5 2>&1 | tail -1 | python -c "$show"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sk_count_stream" -s 2 -c 1 -o gpurun_out/prof_stream_c2_r02d -f python tools/prof_driver.py c2 count 4 > gpurun_out/p2.log 2>&1; tail -2 gpurun_out/p2.log
