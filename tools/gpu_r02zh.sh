# round 2, session zh: C4's mixed-batch kernel, persistent CTAs against short-lived ones (SIB_QM_ROUNDS tiles per CTA)
mkdir -p gpurun_out
for r in 0 1 4 8 32; do
  SIB_QM_ROUNDS=$r timeout 600 python bench.py --no-cpu-baseline --c5-full-intervals 0 --bed-lines 0 --setop-intervals 0 --steps 3 --e2e-steps 1 > gpurun_out/zh_$r.json 2> gpurun_out/zh_$r.err
  python - $r <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/zh_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
c=d['configs']['c4']
print('rounds', sys.argv[1], 'c4 ms', c.get('ms_per_step'), 'mismatches', c.get('parity',{}).get('mismatches'), c.get('error'))
PY
done
( timeout 900 python -m pytest tests/test_genome.py tests/test_gpu_pair.py -m gpu -x -q ) 2>&1 | tail -n 3
