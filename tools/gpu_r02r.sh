# round 2, eighteenth GPU session: resident kernel with a single poll record
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_abi.py -m gpu -q -x -k "single_query or resident" ) > gpurun_out/r02r_tests_a.log 2>&1; echo "resident tests rc=$?"
tail -4 gpurun_out/r02r_tests_a.log
timeout 600 python bench.py --no-cpu-baseline --no-configs --no-sorted --bed-lines 0 --setop-intervals 0 --e2e-steps 1 2>/dev/null | tail -1 > gpurun_out/r02r_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_bench.json').read())
print('latency', d['latency'])
PY
( time timeout 600 python -m pytest tests/test_cpp_header.py tests/test_example_c.py tests/test_gpu_abi.py -m gpu -q -x ) > gpurun_out/r02r_tests_b.log 2>&1; echo "abi tests rc=$?"; tail -3 gpurun_out/r02r_tests_b.log
