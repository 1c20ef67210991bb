#!/bin/bash
# 8-GPU bench line (chained TMEM kernels, fused NVLink all-reduce, graph-captured data-parallel step)
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 \
   --skip-workloads --skip-fp32 > $O/s8_bench$N.log 2>&1; echo "bench$N rc=$?"
grep '^{' $O/s8_bench$N.log | tail -1 > $O/s8_bench$N.json
python - $N <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/s8_bench{sys.argv[1]}.json'))
for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'clocks', 'e2e', 'e2e_philox', 'small_m', 'mid_m'):
    print(k, str(d.get(k))[:200])
print('mc', d.get('mc', {}).get('value'))
PY
tail -3 $O/s8_bench$N.log | cut -c1-300
