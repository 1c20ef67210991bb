#!/bin/bash
# GPU call AB: batched weight-gradient launch at small batches: parity, M = 100 launch list, small-M points, NAIS workloads
mkdir -p gpurun_out; export FBSNN_GBATCH=1
O=gpurun_out
( timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_properties_gpu.py tests/test_round2_gpu.py tests/test_gemm_gpu.py -m gpu -q -x ) > $O/ab_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/ab_pytest.log
( timeout 900 python -m pytest tests/test_chain_gpu.py -m gpu -q -x -k "dispatch or reference" ) > $O/ab_pytest2.log 2>&1; echo "pytest2 rc=$?"
tail -3 $O/ab_pytest2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/ab_fc_m100.csv \
  python bench.py --paths 100 --steps 2 --warmup 3 --skip-mc --skip-cpu --skip-small --skip-variants --skip-e2e --skip-workloads --skip-peak > $O/ab_fc_m100.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ab_fc_m100.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
ks=[(r[ix['Kernel Name']].split('(')[0][:80], float(r[ix['Metric Value']].replace(',',''))) for r in rows[1:]]
last=ks[-15:]
print(f"== FC M=100: last 15 launches {sum(t for _,t in last)/1e3:.1f} us")
for n,t in last: print(f"   {t/1e3:8.1f} us  {n}")
PY
timeout 900 python bench.py --steps 5 --warmup 3 --skip-mc --skip-cpu --skip-variants --skip-e2e --skip-peak > $O/ab_bench.log 2>&1; echo "bench rc=$?"
grep '^{' $O/ab_bench.log | tail -1 > $O/ab_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/ab_bench.json'))
print('value', d['value'], 'small_m', d['small_m']['value'], d['small_m']['ms_per_step'], d['small_m']['gpu_launches'], 'mid_m', d['mid_m']['value'])
for w in d.get('workloads', []):
    if w.get('paths', 0) <= 100: print(w.get('kind'), w.get('dim'), w.get('act'), w.get('paths'), round(w.get('iters_per_s', 0), 1), round(w.get('ms_per_step', 0), 3), w.get('gpu_launches'), w.get('error'))
PY
