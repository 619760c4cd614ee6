# round 2, eighth GPU session: new defaults (no persisting window, cells gathered from HBM), scratch pool, mixed-batch kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"
tail -14 gpurun_out/r02h_tests.log
echo "== persist default"; timeout 300 python tools/exp_r02g.py persist 2>&1 | tail -1 | tee gpurun_out/r02h_persist.json
echo "== build"; timeout 100 python tools/exp_r02g.py build | tail -1
( time timeout 900 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err ); echo "bench rc=$?"
tail -3 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f | sorted %.3f ms frac %.3f | e2e %.2f ms u32 %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['sorted']['ms_per_step'], d['sorted']['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['u32_counts']['ms_per_step']))
print('build', d['build']['ms'], 'sv', d['search_values']['ms_per_step'], d['search_values']['kernel_ms_per_step'], 'sv e2e', d['search_values']['e2e']['ms_per_step'])
c=d['configs']; print('c1', c['c1']['count']['ms_per_step'], 'c5', c['c5_lite']['count']['ms_per_step'], c['c5_lite']['build'], 'c4', c['c4']['ms_per_step'], c['c4']['value']/1e9, c['c4']['parity'])
print('latency', d['latency']['countOverlaps_us'], d['latency']['searchValues_us'], 'bed', d['bed_ingest']['value']/1e6, d['bed_ingest']['cpu_baseline'])
print('wall', d['wall_s'])
PY
echo "== c5 full"; ( time timeout 600 python tools/run_configs.py c5 ) 2>&1 | tail -6 | tee gpurun_out/r02h_c5_full.log
