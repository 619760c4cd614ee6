# round 2, fifteenth GPU session: full suite on the head (done-flag single queries, BED name check), latency and the bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"
tail -14 gpurun_out/r02o_tests.log
( time timeout 900 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err ); echo "bench rc=$?"
tail -3 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02o_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f | sorted %.3f ms | e2e %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['sorted']['ms_per_step'], d['e2e']['ms_per_step']))
print('build', d['build']['ms'], 'sv', d['search_values']['ms_per_step'], d['search_values']['roofline'].get('l2'))
print('latency', d['latency'])
print('bed', d['bed_ingest']['value']/1e6, d['bed_ingest']['grouped_by_contig'])
print('wall', d['wall_s'])
PY
