# round 2, last capture: ncu --set full of the streaming count kernel at 3 CTAs/SM (C2 sorted)
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -f -k regex:sk_count_stream_kernel -s 2 -c 1 -o gpurun_out/prof_stream_c2_r02zv python tools/prof_driver.py c2 count 4 > gpurun_out/zv.log 2>&1; tail -n 1 gpurun_out/zv.log
