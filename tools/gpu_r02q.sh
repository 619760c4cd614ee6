# round 2, seventeenth GPU session: resident single-query kernel (tests under a timeout, then latency)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_abi.py -m gpu -q -x -k "single_query or resident" ) > gpurun_out/r02q_tests_a.log 2>&1; echo "resident tests rc=$?"
tail -6 gpurun_out/r02q_tests_a.log
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r02q_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-configs --no-sorted --bed-lines 0 --setop-intervals 0 --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/r02q_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read())
print('latency', d['latency'])
PY
