# round 2, final session (1 GPU): full GPU suite, smoke, the default bench line, the reference arm, the launch list of the bench command
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02zt_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r02zt_tests.log | tail -n 2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
( time timeout 900 python bench.py > gpurun_out/r02zt_bench.json 2> gpurun_out/r02zt_bench.err ); echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02zt_ref.json 2> gpurun_out/r02zt_ref.err; tail -c 600 gpurun_out/r02zt_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02zt.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --bed-lines 0 --setop-intervals 0 --c4-scale 0.1 --c5-full-intervals 0 > gpurun_out/zt_launches.log 2>&1; tail -c 300 gpurun_out/zt_launches.log
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02zt_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.4f frac_l2 %.3f | sorted %.3f ms %.3f | e2e %.2f ms ceiling %.2f frac %.3f | u32 %.2f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['l2']['frac_l2'], d['sorted']['ms_per_step'], d['sorted']['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['copy_ceiling_ms'], d['e2e']['frac_of_copy_ceiling'], d['e2e']['u32_counts']['ms_per_step']))
sv=d['search_values']; print('sv', sv['ms_per_step'], sv['value']/1e9, sv['kernel_ms_per_step'], 'e2e', sv['e2e']['ms_per_step'])
c=d['configs']; print('c1', c['c1']['count']['ms_per_step'], c['c1']['search_values']['ms_per_step'], 'c5lite', c['c5_lite']['count']['ms_per_step'], 'c5', c['c5']['count']['ms_per_step'], c['c5']['build']['ms'], c['c5']['hbm_sector_gather']['frac'], 'c4', c['c4'].get('ms_per_step'), c['c4'].get('error'))
print('build', d['build']['ms'], 'latency', d['latency']['resident'], 'cpu', d['cpu_baseline']['value'])
print('wall', d['wall_s'])
PY
