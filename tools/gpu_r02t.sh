# round 2, twentieth GPU session: default bench line with the full-size C5 arm
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err ); echo "bench rc=$?"
tail -3 gpurun_out/r02t_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02t_bench.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f' % (d['value']/1e9, d['ms_per_step']))
print('c5', json.dumps(d['configs'].get('c5'))[:1500])
print('wall', d['wall_s'])
PY
