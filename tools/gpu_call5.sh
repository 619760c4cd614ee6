# experiment call: fill with 8 CTAs/SM; count-cells LDGSTS variants
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stab or rank_cells" > gpurun_out/tests_fill.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_fill.log; tail -3 gpurun_out/tests_fill.log
timeout 600 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 600 gpurun_out/bench_quick.json
timeout 900 bash tools/variants.sh --steps 5 > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
