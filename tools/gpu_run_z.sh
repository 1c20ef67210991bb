#!/bin/bash
# GPU call Z: path_advance with prefetched increments: parity (X bit-identical), M = 100 launch list, small-M bench points
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_properties_gpu.py tests/test_round2_gpu.py -m gpu -q -x ) > $O/z_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/z_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/z_fc_m100.csv \
  python bench.py --paths 100 --steps 2 --warmup 3 --skip-mc --skip-cpu --skip-small --skip-variants --skip-e2e --skip-workloads --skip-peak > $O/z_fc_m100.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/z_fc_m100.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
ks=[(r[ix['Kernel Name']].split('(')[0][:80], float(r[ix['Metric Value']].replace(',',''))) for r in rows[1:]]
last=ks[-18:]
tot=sum(t for _,t in last)
print(f"== FC M=100: last 18 launches {tot/1e3:.1f} us")
for n,t in last: print(f"   {t/1e3:8.1f} us  {n}")
PY
timeout 600 python bench.py --steps 5 --warmup 3 --skip-mc --skip-cpu --skip-variants --skip-e2e --skip-peak --skip-workloads > $O/z_bench.log 2>&1
grep '^{' $O/z_bench.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'small_m', d['small_m']['value'], d['small_m']['ms_per_step'], 'mid_m', d['mid_m']['value'])"
