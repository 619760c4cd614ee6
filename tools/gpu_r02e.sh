# round 2, fifth GPU session: full GPU suite on the head (malformed side list included)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 ) > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r02e_tests.log
