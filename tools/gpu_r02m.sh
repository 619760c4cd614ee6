# round 2, thirteenth GPU session (2 GPUs): fused count + all-gather over peer memory (multi.cu in one process; CUDA IPC under torchrun)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_genome.py -m gpu -q -x --durations=5 ) > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r02m_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err ); echo "bench n2 rc=$?"
tail -5 gpurun_out/r02m_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02m_bench_n2.json') if l.startswith('{')][-1])
print('value %.1f G ms %.3f' % (d['value']/1e9, d['ms_per_step']))
print('strong', json.dumps(d.get('strong'))[:2500])
c=d.get('configs') or {}
print('c4', json.dumps(c.get('c4'))[:2000])
print('wall', d['wall_s'])
PY
echo "== multi.cu: NCCL gather vs fused"
SIB_MULTI_P2P=0 timeout 300 python - <<'PY'
import numpy as np, time
from superintervals_b200 import workloads as W
from superintervals_b200.multi import MultiIndex
s,e,qs,qe = W.config2(10_000_000, 50_000_000, 2)
m = MultiIndex(); m.build(s,e)
for _ in range(3): c = m.count_batch(qs,qe)
print('nccl ', m.stats())
PY
timeout 300 python - <<'PY'
import numpy as np, time
from superintervals_b200 import workloads as W
from superintervals_b200.multi import MultiIndex
s,e,qs,qe = W.config2(10_000_000, 50_000_000, 2)
m = MultiIndex(); m.build(s,e)
for _ in range(3): c = m.count_batch(qs,qe)
print('fused', m.stats())
PY
