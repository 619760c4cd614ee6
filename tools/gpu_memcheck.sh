# compute-sanitizer memcheck over the GPU parity suites (usage: bash tools/gpu_memcheck.sh [pytest args])
mkdir -p gpurun_out
timeout 540 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest -q -x -p no:cacheprovider \
  ${@:-tests/test_gpu_parity.py tests/test_gpu_setops.py tests/test_gpu_bed.py tests/test_gpu_abi.py} -m gpu > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|rc=" gpurun_out/memcheck.log | tail -15
