# round 2, nineteenth GPU session: resident single queries as the default -- whole suite, smoke, bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 ) > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r02s_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err ); echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02s_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f | sorted %.3f ms | e2e %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['sorted']['ms_per_step'], d['e2e']['ms_per_step']))
print('latency', d['latency'])
print('wall', d['wall_s'])
PY
