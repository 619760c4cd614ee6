# round-1 session-6 GPU call: parity tests, bench with the run+stab fill, ncu evidence of search_values on C3 (r01e)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log; tail -5 gpurun_out/tests.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_scan|qk_fill|qk_count" -c 6 -o gpurun_out/prof_search_c3_r01e2 -f python tools/prof_driver.py c3 search 1 > gpurun_out/p3.log 2>&1
ls -la gpurun_out
