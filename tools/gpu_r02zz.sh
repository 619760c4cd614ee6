# round 2, last call: compute-sanitizer memcheck over the streaming kernel (3 CTAs/SM) and build() (ballot ranking)
mkdir -p gpurun_out
timeout 90 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -m gpu -q -x -k "stream or build" > gpurun_out/memcheck_r02zz.txt 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02zz.txt | tail -n 3
