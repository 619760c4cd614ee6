# round 2, session zg: HBM gather microbench in the count kernels' own launch shape; ncu --set full of the paired cells kernel on C5-lite
mkdir -p gpurun_out
timeout 600 tools/bin/hbm_gather > gpurun_out/hbm_gather3.json 2> gpurun_out/hbm_gather3.err; cat gpurun_out/hbm_gather3.json
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:qk_count_cells_kernel -s 2 -c 1 -o gpurun_out/prof_pair_c5lite_r02zg python tools/prof_driver.py c5lite count 4 > gpurun_out/zg_ncu.log 2>&1; tail -n 1 gpurun_out/zg_ncu.log
