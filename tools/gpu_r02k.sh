# round 2, eleventh GPU session: ncu --set full captures of the hot kernels on the final defaults + launch list of the bench command
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02k_tests.log
N="ncu --set full --clock-control none --import-source on -f"
timeout 600 $N -k regex:qk_count_cells_kernel -s 2 -c 1 -o gpurun_out/prof_cells_c2_r02k python tools/prof_driver.py c2 count_unsorted 4 > gpurun_out/p1.log 2>&1; tail -1 gpurun_out/p1.log
timeout 600 $N -k regex:sk_count_stream_kernel -s 2 -c 1 -o gpurun_out/prof_stream_c2_r02k python tools/prof_driver.py c2 count 4 > gpurun_out/p2.log 2>&1; tail -1 gpurun_out/p2.log
timeout 600 $N -k regex:qk_fill_runs_kernel -s 2 -c 1 -o gpurun_out/prof_fill_c3_r02k python tools/prof_driver.py c3 search 4 > gpurun_out/p3.log 2>&1; tail -1 gpurun_out/p3.log
timeout 600 $N -k regex:qk_count_mixed_kernel -s 1 -c 1 -o gpurun_out/prof_mixed_r02k python tools/prof_driver.py mixed x 3 > gpurun_out/p4.log 2>&1; tail -1 gpurun_out/p4.log
timeout 600 $N -k regex:rs_onesweep_kernel -s 9 -c 2 -o gpurun_out/prof_sort_c2_r02k python tools/exp_r02g.py build > gpurun_out/p5.log 2>&1; tail -1 gpurun_out/p5.log
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --bed-lines 0 --setop-intervals 0 --c4-scale 0.1 > gpurun_out/r02k_launch_bench.log 2>&1; tail -c 300 gpurun_out/r02k_launch_bench.log; wc -l gpurun_out/launches_r02k.csv
echo "== bench (clean)"
( time timeout 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err ); echo "bench rc=$?"
ls -la gpurun_out/*r02k*
