# round 2, ninth GPU session: narrow sort, persistent mixed-batch kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"
tail -14 gpurun_out/r02i_tests.log
echo "== build"; timeout 100 python tools/exp_r02g.py build | tail -1
( time timeout 900 python bench.py --no-cpu-baseline --bed-lines 0 --setop-intervals 0 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err ); echo "bench rc=$?"
tail -3 gpurun_out/r02i_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f | sorted %.3f ms frac %.3f | e2e %.2f ms u32 %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['sorted']['ms_per_step'], d['sorted']['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['u32_counts']['ms_per_step']))
print('build', d['build'])
c=d['configs']; print('c1', c['c1']['count']['ms_per_step'], 'c5', c['c5_lite']['count']['ms_per_step'], c['c5_lite']['build'], 'c4', c['c4']['ms_per_step'], c['c4']['value']/1e9, c['c4']['parity'], c['c4']['build_ms_max_over_ranks'])
print('wall', d['wall_s'])
PY
echo "== build launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02i_build_launches.csv python tools/exp_r02g.py build > gpurun_out/r02i_build.log 2>&1; tail -1 gpurun_out/r02i_build.log
echo "== c5 full"; ( time timeout 600 python tools/run_configs.py c5 ) 2>&1 | tail -6 | tee gpurun_out/r02i_c5_full.log
