# ncu evidence for profiles/: launch list of one bench run + full captures of the hot kernels.
set -x
ncu --set full --clock-control none --import-source on -k regex:"pt_onesweep|qk_count_rank" -s 3 -c 3 -o gpurun_out/prof_partition_c2 -f python tools/prof_driver.py c2 count_unsorted 2 > gpurun_out/p1.log 2>&1
ls -la gpurun_out
