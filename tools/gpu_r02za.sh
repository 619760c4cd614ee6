# round 2, session za: (1) HBM gather microbench: single sector vs adjacent pair vs line; (2) DRAM bytes of the cells kernel with and without L2 eviction hints
mkdir -p gpurun_out
timeout 300 tools/bin/hbm_gather > gpurun_out/hbm_gather.json 2> gpurun_out/hbm_gather.err; cat gpurun_out/hbm_gather.json
M="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --csv"
timeout 400 $M -k regex:qk_count_cells_kernel -s 2 -c 2 --log-file gpurun_out/r02za_cells_default.csv python tools/prof_driver.py c2 count_unsorted 4 > gpurun_out/p1.log 2>&1
SIB_LIBRARY=$PWD/superintervals_b200/variants/lib_qc_hints.so timeout 400 $M -k regex:qk_count_cells_kernel -s 2 -c 2 --log-file gpurun_out/r02za_cells_hints.csv python tools/prof_driver.py c2 count_unsorted 4 > gpurun_out/p2.log 2>&1
tail -4 gpurun_out/r02za_cells_default.csv gpurun_out/r02za_cells_hints.csv
