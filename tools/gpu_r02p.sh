# round 2, sixteenth GPU session: 128-wide single search, device-side widening of coverage counts; suite + latency
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r02p_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-configs --no-sorted --bed-lines 0 --setop-intervals 0 --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/r02p_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench.json').read())
print('latency', d['latency'])
PY
