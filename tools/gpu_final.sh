# round-1 final evidence call: full GPU suite, smoke, bench (both arms), launch list, ncu captures (r01f)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log; tail -16 gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1500 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_count_cells" -s 2 -c 1 -o gpurun_out/prof_cells_c2_r01f -f python tools/prof_driver.py c2 count_unsorted 4 > gpurun_out/p1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qk_scan|qk_fill|qk_count" -c 6 -o gpurun_out/prof_search_c3_r01f -f python tools/prof_driver.py c3 search 1 > gpurun_out/p3.log 2>&1
ls -la gpurun_out
