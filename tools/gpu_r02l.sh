# round 2, twelfth GPU session: fan-out / barrier tests on one GPU, tuning variants of the mixed-batch kernel and the radix sort
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_device.py tests/test_genome.py tests/test_gpu_multi.py tests/test_gpu_stream.py -m gpu -q -x ) > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02l_tests.log
echo "== mixed default"; timeout 200 python tools/exp_r02g.py mixed | tail -1
for f in superintervals_b200/variants/lib_qm*.so; do echo "== $f"; SIB_LIBRARY=$PWD/$f timeout 200 python tools/exp_r02g.py mixed | tail -1; done
echo "== build default"; timeout 100 python tools/exp_r02g.py build | tail -1
for f in superintervals_b200/variants/lib_rs_*.so; do echo "== $f"; SIB_LIBRARY=$PWD/$f timeout 100 python tools/exp_r02g.py build | tail -1; done
