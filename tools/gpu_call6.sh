# set-algebra parity on the GPU + full suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_setops.py -m gpu -x -q > gpurun_out/tests_setops.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_setops.log; tail -30 gpurun_out/tests_setops.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log; tail -5 gpurun_out/tests.log
