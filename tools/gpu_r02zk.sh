# round 2, session zk: host-batch pipeline with 2 M-query chunks on 4 slots; L2 gather peak in the count kernels' launch shape; full suite
mkdir -p gpurun_out
timeout 600 tools/bin/l2_peak > gpurun_out/l2_peak2.json 2> gpurun_out/l2_peak2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/l2_peak2.json'))
for r in d['gather']:
    if r['ctas_per_sm']==2: print(r['table_mb'], 'flat %.4g k8 %.4g' % (r['sectors_per_s_flat'], r['sectors_per_s_k8']))
print(d['stream'])
PY
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02zk_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/r02zk_tests.log
( time timeout 900 python bench.py > gpurun_out/r02zk_bench.json 2> gpurun_out/r02zk_bench.err ); echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02zk_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.4f frac_l2 %.3f | sorted %.3f ms %.3f | e2e %.2f ms ceiling %.2f frac %.3f | u32 %.2f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['l2']['frac_l2'], d['sorted']['ms_per_step'], d['sorted']['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['copy_ceiling_ms'], d['e2e']['frac_of_copy_ceiling'], d['e2e']['u32_counts']['ms_per_step']))
print('sorted e2e', d['sorted']['e2e']['ms_per_step'])
sv=d['search_values']; print('sv', sv['ms_per_step'], sv['value']/1e9, sv['kernel_ms_per_step'], 'e2e', sv['e2e']['ms_per_step'])
c=d['configs']; print('c1', c['c1']['count']['ms_per_step'], c['c1']['search_values']['ms_per_step'], 'c5lite', c['c5_lite']['count']['ms_per_step'], 'c5', c['c5']['count']['ms_per_step'], c['c5']['build']['ms'], c['c5']['hbm_sector_gather']['frac'], 'c4', c['c4'].get('ms_per_step'), c['c4'].get('error'))
print('build', d['build']['ms'], 'latency', d['latency']['resident'])
print('wall', d['wall_s'])
PY
