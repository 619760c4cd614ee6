# round 2, fourteenth GPU session (8 GPUs): the bench line as the driver's scaling run launches it, N = 8 and N = 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02n_bench_n$N.json 2> gpurun_out/r02n_bench_n$N.err ); echo "bench n$N rc=$?"
  tail -3 gpurun_out/r02n_bench_n$N.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench_n$N.json') if l.startswith('{')][-1])
print('N=$N value %.1f G ms %.3f | e2e %.2f G (%.1f ms, ceiling %.1f ms)' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['e2e'].get('copy_ceiling_ms') or 0))
st=d.get('strong') or {}
print('strong value %.1f G (%s) nccl %.1f G count_only %.1f G fused %s' % (st.get('value',0)/1e9, st.get('value_is'), st.get('value_nccl',0)/1e9, st.get('value_count_only',0)/1e9, json.dumps(st.get('fused'))[:400]))
c=(d.get('configs') or {}).get('c4') or {}
print('c4 part %.1f ms %.1f G | repl %s' % (c.get('ms_per_step',0), c.get('value',0)/1e9, json.dumps(c.get('replicated'))[:300]), c.get('parity'))
print('sv', json.dumps(d.get('search_values'))[:400])
print('wall', d['wall_s'])
PY
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02n_ref_n8.json 2> gpurun_out/r02n_ref_n8.err ); echo "ref n8 rc=$?"
tail -c 300 gpurun_out/r02n_ref_n8.json
