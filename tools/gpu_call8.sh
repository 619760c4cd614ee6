set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_python_setops.py tests/test_gpu_bed.py tests/test_cpp_header.py -m gpu -q --durations=5 > gpurun_out/tests_new2.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_new2.log; tail -40 gpurun_out/tests_new2.log
timeout 600 python tools/search_dense.py 20000000 > gpurun_out/search_dense.json 2> gpurun_out/search_dense.err; cat gpurun_out/search_dense.json; tail -5 gpurun_out/search_dense.err
