#!/bin/bash
# GPU call T (2 GPUs): multi-GPU equivalence test + 2-GPU bench line with the chained TMEM kernels as default
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > $O/t_pytest2.log 2>&1; echo "pytest2 rc=$?"
tail -3 $O/t_pytest2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
   --skip-workloads --skip-fp32 > $O/t_bench2.log 2>&1; echo "bench2 rc=$?"
grep '^{' $O/t_bench2.log | tail -1 > $O/t_bench2.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/t_bench2.json'))
for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'clocks', 'e2e', 'e2e_philox', 'small_m', 'mid_m', 'collective'):
    print(k, str(d.get(k))[:300])
print('mc', d.get('mc', {}).get('value'))
PY
tail -5 $O/t_bench2.log | cut -c1-300
