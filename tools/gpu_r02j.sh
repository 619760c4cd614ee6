# round 2, tenth GPU session (2 GPUs): multi-device C entry points over NCCL, bench under torchrun (strong scaling + all-gather, C4 all-to-all)
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_genome.py tests/test_gpu_device.py -m gpu -q -x --durations=5 ) > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r02j_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02j_bench_n2.json 2> gpurun_out/r02j_bench_n2.err ); echo "bench n2 rc=$?"
tail -5 gpurun_out/r02j_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02j_bench_n2.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f | e2e %s' % (d['value']/1e9, d['ms_per_step'], json.dumps(d['e2e'])[:400]))
print('strong', json.dumps(d.get('strong'))[:1500])
c=d.get('configs') or {}
print('c4', json.dumps(c.get('c4'))[:1200])
print('sv', json.dumps(d.get('search_values'))[:800])
print('wall', d['wall_s'])
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02j_ref_n2.json 2> gpurun_out/r02j_ref_n2.err ); echo "ref n2 rc=$?"
tail -c 400 gpurun_out/r02j_ref_n2.json
