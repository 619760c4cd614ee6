# round 2, session zd: pair cells -- parity, then C5-lite / C5-like counts with and without them
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -x -q ) 2>&1 | tail -n 15
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py tests/test_genome.py tests/test_gpu_stream.py -m gpu -x -q ) 2>&1 | tail -n 3
timeout 600 python tools/pair_probe.py 2>&1 | tail -n 12
