set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qk_count -s 1 -c 1 -o gpurun_out/prof_count_c2_r01b -f python tools/prof_driver.py c2 count_unsorted 2 > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qk_count -s 1 -c 1 -o gpurun_out/prof_count_c2sorted_r01b -f python tools/prof_driver.py c2 count 2 > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qk_fill|qk_scan|qk_count" -s 3 -c 3 -o gpurun_out/prof_search_c3_r01b -f python tools/prof_driver.py c3 search 2 > gpurun_out/p3.log 2>&1
python bench.py --order sorted --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_sorted.json
ls -la gpurun_out
