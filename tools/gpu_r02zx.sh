# round 2, last call: compute-sanitizer memcheck over the pair-cells kernels (two formats, over-full sides)
mkdir -p gpurun_out
timeout 130 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pair.py -m gpu -q -x -k "dense_c5 or read_c2 or overfull" > gpurun_out/memcheck_r02zx.txt 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02zx.txt | tail -n 3
