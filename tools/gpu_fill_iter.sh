# pooled stab candidates: parity, C3 search bench, dense search, ncu of the fill
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py tests/test_gpu_abi.py -m gpu -x -q > gpurun_out/tests_fill.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_fill.log; tail -5 gpurun_out/tests_fill.log
timeout 600 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 300 gpurun_out/bench_quick.json
timeout 600 python tools/search_dense.py 20000000 > gpurun_out/search_dense.json 2> gpurun_out/search_dense.err; cat gpurun_out/search_dense.json; tail -3 gpurun_out/search_dense.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_fill" -c 2 -o gpurun_out/prof_fill_c3_r01f -f python tools/prof_driver.py c3 search 2 > gpurun_out/p5.log 2>&1
