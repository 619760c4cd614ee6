# round 2, last call: compute-sanitizer memcheck over the rest of the pair-cells tests and the mixed-batch kernel
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pair.py tests/test_genome.py -m gpu -q -x -k "not dense_c5 and not read_c2 and not overfull" > gpurun_out/memcheck_r02zy.txt 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02zy.txt | tail -n 3
