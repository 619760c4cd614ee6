# round 2, session ze: suite + bench line with pair cells (C4, C5, C5-lite) and the ballot ranking of the radix sort
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02ze_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02ze_tests.log
( time timeout 900 python bench.py > gpurun_out/r02ze_bench.json 2> gpurun_out/r02ze_bench.err ); echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02ze_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.4f frac_l2 %.3f | sorted %.3f ms %.3f | e2e %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['roofline']['l2']['frac_l2'], d['sorted']['ms_per_step'], d['sorted']['roofline']['frac'], d['e2e']['ms_per_step']))
sv=d['search_values']; print('sv', sv['ms_per_step'], sv['value']/1e9, sv['kernel_ms_per_step'], sv['roofline'].get('l2',{}).get('frac_l2'), 'e2e', sv['e2e']['ms_per_step'])
c=d['configs']; print('c1', c['c1']['count']['ms_per_step'], c['c1']['search_values']['ms_per_step'], 'c5lite', c['c5_lite']['count']['ms_per_step'], 'c5', c['c5']['count']['ms_per_step'], c['c5']['build']['ms'], 'c4', c['c4']['ms_per_step'])
print('build', d['build']['ms'], 'latency', d['latency']['resident'])
print('wall', d['wall_s'])
PY
