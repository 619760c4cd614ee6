# round 2, twenty-fourth GPU session: ncu --set full of the two kernels that changed (128-thread cells CTAs, value lists in the fill)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 600 $N -k regex:qk_count_cells_kernel -s 2 -c 1 -o gpurun_out/prof_cells_c2_r02x python tools/prof_driver.py c2 count_unsorted 4 > gpurun_out/p1.log 2>&1; tail -1 gpurun_out/p1.log
timeout 600 $N -k regex:qk_fill_runs_kernel -s 2 -c 1 -o gpurun_out/prof_fill_c3_r02x python tools/prof_driver.py c3 search 4 > gpurun_out/p3.log 2>&1; tail -1 gpurun_out/p3.log
ls -la gpurun_out/*r02x*
