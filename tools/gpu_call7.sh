# new rows on the GPU: set algebra (C ABI, C++, Python), BED ingest, BED throughput
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setops.py tests/test_gpu_python_setops.py tests/test_gpu_bed.py tests/test_cpp_header.py -m gpu -x -q --durations=8 > gpurun_out/tests_new.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_new.log; tail -40 gpurun_out/tests_new.log
timeout 600 python tools/bed_bench.py 50000000 5000000 > gpurun_out/bed_bench.json 2> gpurun_out/bed_bench.err; cat gpurun_out/bed_bench.json; tail -5 gpurun_out/bed_bench.err
