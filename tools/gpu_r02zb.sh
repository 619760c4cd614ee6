# round 2, session zb: radix ranking by ballots vs MATCH.ANY; build breakdown at 250 M intervals
mkdir -p gpurun_out
for n in 1e7 1e8; do
  echo "== ballots n=$n"; timeout 300 python tools/build_probe.py $n 3 c5 2>&1 | tail -1
  echo "== match   n=$n"; SIB_LIBRARY=$PWD/superintervals_b200/variants/lib_rs_match.so timeout 300 python tools/build_probe.py $n 3 c5 2>&1 | tail -1
done
echo "== c2 law 1e7 ballots"; timeout 300 python tools/build_probe.py 1e7 3 c2 2>&1 | tail -1
echo "== c2 law 1e7 match"; SIB_LIBRARY=$PWD/superintervals_b200/variants/lib_rs_match.so timeout 300 python tools/build_probe.py 1e7 3 c2 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02zb_build250m_launches.csv python tools/build_probe.py 2.5e8 1 c5 > gpurun_out/zb_ncu.log 2>&1; tail -n 1 gpurun_out/zb_ncu.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py -m gpu -x -q ) 2>&1 | tail -n 3
