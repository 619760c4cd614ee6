# round 2, session zf: HBM gather microbench with K = 1..8 and the L2 fetch granularity hint; pair-cells tests; C4 of the bench; search_values e2e
mkdir -p gpurun_out
timeout 600 tools/bin/hbm_gather > gpurun_out/hbm_gather2.json 2> gpurun_out/hbm_gather2.err; cat gpurun_out/hbm_gather2.json
( timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_abi.py -m gpu -x -q ) 2>&1 | tail -n 3
( time timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02zf_bench.json 2> gpurun_out/r02zf_bench.err ); echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02zf_bench.json') if l.startswith('{')][-1])
c=d['configs']; print('c4', json.dumps(c['c4'])[:1500])
print('sv e2e', d['search_values']['e2e'])
PY
