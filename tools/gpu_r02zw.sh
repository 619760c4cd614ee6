# round 2, last 2-GPU check: the bench line under torchrun on the head
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02zw_bench_n2.json 2> gpurun_out/r02zw_bench_n2.err ); echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02zw_bench_n2.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f e2e %.2f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step']))
c=d['configs']['c4']; print('c4', c.get('ms_per_step'), c.get('value_is'), (c.get('peer') or {}).get('ms_per_step'), (c.get('peer') or {}).get('equals_dispatched_counts'), c.get('error'))
PY
