# round 2, third GPU session: parity suites on the new kernels, L2 persistence, stream/cells variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py tests/test_gpu_device.py tests/test_cpp_header.py tests/test_gpu_abi.py -q > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02c_tests.log
show='import json,sys
d=json.loads(sys.stdin.read()); k=d["kernels"]
print("  step %.3f ms  %.2f Gq/s | " % (d["ms_per_step"], d["value"]/1e9) + "  ".join("%s %.3f" % (n, v["ms_per_launch"]) for n, v in k.items()) + " | mismatches %d e2e %.2f ms" % (d["parity"]["mismatches"], d["e2e"]["ms_per_step"]))'
echo "== sorted (stream, default)"; timeout 300 python bench.py --order sorted --no-cpu-baseline --no-search-values --e2e-steps 2 --steps 10 2> gpurun_out/r02c_sorted.err | tail -1 | tee gpurun_out/r02c_sorted.json | python -c "$show"
echo "== shuffled (default pt1mb8, L2 persist)"; timeout 300 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 2 --steps 10 2> gpurun_out/r02c_shuf.err | tail -1 | tee gpurun_out/r02c_shuf.json | python -c "$show"
echo "== shuffled (default, SIB_L2_PERSIST=0)"; SIB_L2_PERSIST=0 timeout 300 python bench.py --no-cpu-baseline --no-search-values --e2e-steps 1 --steps 10 2>&1 | tail -1 | python -c "$show"
for f in superintervals_b200/variants/lib_*.so; do
  case $f in *lib_sk*) ord="--order sorted";; *) ord="";; esac
  echo "== $f $ord"; SIB_LIBRARY=$PWD/$f timeout 200 python bench.py $ord --no-cpu-baseline --no-search-values --e2e-steps 1 --steps 5 2>&1 | tail -1 | python -c "$show"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_count_cells" -s 2 -c 1 -o gpurun_out/prof_cells_c2_r02c -f python tools/prof_driver.py c2 count_unsorted 4 > gpurun_out/p1.log 2>&1; tail -2 gpurun_out/p1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sk_count_stream" -s 2 -c 1 -o gpurun_out/prof_stream_c2_r02c -f python tools/prof_driver.py c2 count 4 > gpurun_out/p2.log 2>&1; tail -2 gpurun_out/p2.log
