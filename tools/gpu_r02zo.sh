# round 2, session zo (2 GPUs): the two-process tests only
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/r02zo_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/r02zo_tests.log
