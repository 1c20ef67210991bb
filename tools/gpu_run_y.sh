#!/bin/bash
# GPU call Y: NAIS-Net projection kernels on a grid + batched split-K reductions: parity tests, then workloads of the bench
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_properties_gpu.py tests/test_round2_gpu.py -m gpu -q -x ) > $O/y_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/y_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --skip-mc --skip-cpu --skip-small --skip-variants --skip-e2e --skip-peak > $O/y_bench.log 2>&1; echo "bench rc=$?"
grep '^{' $O/y_bench.log | tail -1 > $O/y_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/y_bench.json'))
for w in d.get('workloads', []):
    print(w.get('kind'), w.get('dim'), w.get('act'), w.get('paths'), round(w.get('iters_per_s', 0), 1), round(w.get('ms_per_step', 0), 3), w.get('gpu_launches'), w.get('error'))
PY
