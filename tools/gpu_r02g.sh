# round 2, seventh GPU session: persisting-L2 window on/off, C5-lite direct vs partition, build launch list
mkdir -p gpurun_out
for p in 1 0; do
  echo "== persist $p"; SIB_L2_PERSIST=$p timeout 300 python tools/exp_r02g.py persist 2>&1 | tail -1 | tee gpurun_out/r02g_persist$p.json
done
echo "== c5"; timeout 300 python tools/exp_r02g.py c5 2>&1 | tail -1 | tee gpurun_out/r02g_c5.json
echo "== c5 128M"; timeout 300 python tools/exp_r02g.py c5 128000000 2>&1 | tail -1 | tee gpurun_out/r02g_c5_128.json
echo "== build launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02g_build_launches.csv python tools/exp_r02g.py build > gpurun_out/r02g_build.log 2>&1; tail -1 gpurun_out/r02g_build.log
timeout 100 python tools/exp_r02g.py build | tail -1
