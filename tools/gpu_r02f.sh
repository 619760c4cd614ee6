# round 2, sixth GPU session: full suite on the head, the default bench line and the reference arm as the driver runs them
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=10 ) > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"
tail -18 gpurun_out/r02f_tests.log
( time timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err ); echo "bench rc=$?"
tail -5 gpurun_out/r02f_bench.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err ); echo "ref rc=$?"
tail -c 600 gpurun_out/r02f_ref.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
