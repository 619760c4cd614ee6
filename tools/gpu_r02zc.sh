# round 2, session zc: ncu --set full of the onesweep passes (ballot ranking) at 100 M intervals
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:rs_onesweep -s 2 -c 4 -o gpurun_out/prof_sort_100m_r02zc python tools/build_probe.py 1e8 1 c5 > gpurun_out/zc_ncu.log 2>&1; tail -n 2 gpurun_out/zc_ncu.log
ls -la gpurun_out/prof_sort_100m_r02zc.ncu-rep
tools/bin/cub_sort 100000000 2>&1 | tail -n 3
tools/bin/cub_sort 250000000 2>&1 | tail -n 3
