# ncu evidence for profiles/: launch list of one bench run + full captures of the hot kernels.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search-values --e2e-steps 1 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pt_onesweep|qk_count_rank|pt_histogram" -s 4 -c 4 -o gpurun_out/prof_count_c2_r01c -f python tools/prof_driver.py c2 count_unsorted 2 > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qk_count_rank" -s 1 -c 1 -o gpurun_out/prof_count_c2sorted_r01c -f python tools/prof_driver.py c2 count 2 > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qk_count_kernel" -s 1 -c 1 -o gpurun_out/prof_countwalk_c2_r01c -f python tools/prof_driver.py c2 count_walk 2 > gpurun_out/p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qk_fill|qk_scan" -s 2 -c 2 -o gpurun_out/prof_search_c3_r01c -f python tools/prof_driver.py c3 search 2 > gpurun_out/p3.log 2>&1
ls -la gpurun_out
