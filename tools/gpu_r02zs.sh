# round 2, session zs (4 GPUs): peer mode with all contig ids in flight: the bench line under torchrun, C4 in its three forms (peer / dispatched / replicated)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "peer or fused" ) 2>&1 | tail -n 2
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02zs_bench_n4.json 2> gpurun_out/r02zs_bench_n4.err ); echo "bench n4 rc=$?"
tail -n 3 gpurun_out/r02zs_bench_n4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02zs_bench_n4.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f e2e %.2f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step']))
st=d['strong']; print('strong', st['value']/1e9, st['value_is'], st['value_nccl']/1e9)
c=d['configs']['c4']; print('c4', c.get('ms_per_step'), c.get('value_is'), 'dispatched', (c.get('dispatched') or {}).get('ms_per_step'), 'replicated', (c.get('replicated') or {}).get('ms_per_step'), 'peer', json.dumps(c.get('peer'))[:400], c.get('error'))
print('wall', d['wall_s'])
PY
