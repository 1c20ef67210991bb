#!/bin/bash
# GPU call J (round 2): state of the tree -- full GPU test-suite, smoke, default bench line, reference arm, ncu launch list.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/j_summary.txt
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > $O/j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/j_summary.txt
timeout 300 python __graft_entry__.py smoke > $O/j_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/j_summary.txt
( time timeout 1500 python bench.py ) > $O/j_bench.log 2>&1; echo "bench rc=$?" | tee -a $O/j_summary.txt
grep '^{' $O/j_bench.log | tail -1 > $O/j_bench.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/j_ref.log 2>&1; echo "ref rc=$?" | tee -a $O/j_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/j_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-mc --skip-cpu --skip-small --skip-variants --skip-e2e --skip-workloads --skip-peak > $O/j_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $O/j_summary.txt
tail -8 $O/j_pytest.log
tail -3 $O/j_smoke.log
cat $O/j_summary.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/j_bench.json'))
for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks', 'e2e', 'e2e_philox', 'small_m', 'mid_m'):
    print(k, d.get(k))
r = d['roofline']
print('roofline', {k: r[k] for k in ('bound', 'achieved', 'peak', 'frac', 'traffic')})
for t in r['launch_table']:
    print(t)
print('variants', d.get('variants'))
for w in d.get('workloads', []):
    print(w.get('kind'), w.get('dim'), w.get('act'), w.get('paths'), w.get('iters_per_s'), w.get('roofline', {}).get('frac'), w.get('error'))
print('mc', {k: v for k, v in d.get('mc', {}).items() if k not in ('config',)})
print('cpu', d.get('cpu_baseline'))
PY
