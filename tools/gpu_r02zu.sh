# round 2, session zu (2 GPUs): remote walks on side streams: contig-partitioned mixed batch read in place over peer memory: suite of the multi-GPU tests and the bench line under torchrun on the head
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_genome.py -m gpu -q -x ) > gpurun_out/r02zu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02zu_tests.log
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02zu_bench_n2.json 2> gpurun_out/r02zu_bench_n2.err ); echo "bench n2 rc=$?"
tail -3 gpurun_out/r02zu_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02zu_bench_n2.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f' % (d['value']/1e9, d['ms_per_step']))
st=d['strong']; print('strong', st['value']/1e9, st['value_is'], st['value_nccl']/1e9)
print('sv', json.dumps(d['search_values'])[:900])
c=d['configs']['c4']; print('c4', c['ms_per_step'], c['value_is'], 'dispatched', c['dispatched'], 'replicated', c['replicated']['ms_per_step'], 'peer', c['peer'])
print('wall', d['wall_s'])
PY
