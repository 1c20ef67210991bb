#!/bin/bash
# GPU call P (round 2): chaint_kernel as default -- full GPU test-suite, smoke, bench, launch list, ncu --set full of the sweeps.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/p_summary.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/p_summary.txt
timeout 300 python __graft_entry__.py smoke > $O/p_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/p_summary.txt
( time timeout 1500 python bench.py ) > $O/p_bench.log 2>&1; echo "bench rc=$?" | tee -a $O/p_summary.txt
grep '^{' $O/p_bench.log | tail -1 > $O/p_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/p_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-mc --skip-cpu --skip-small --skip-variants --skip-e2e --skip-workloads --skip-peak > $O/p_ncu.log 2>&1
echo "ncu-list rc=$?" | tee -a $O/p_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chaint_kernel -s 4 -c 4 -f -o $O/r02_chaint_full_m16384 \
  python tools/one_step.py 16384 tf32x3 > $O/p_ncu_full.log 2>&1
echo "ncu-full rc=$?" | tee -a $O/p_summary.txt
tail -6 $O/p_pytest.log
tail -2 $O/p_smoke.log
cat $O/p_summary.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/p_bench.json'))
for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks', 'e2e', 'e2e_philox', 'small_m', 'mid_m'):
    print(k, d.get(k))
r = d['roofline']
print('roofline', {k: r[k] for k in ('bound', 'achieved', 'peak', 'frac', 'traffic')})
for t in r['launch_table']:
    print(t)
print('variants', d.get('variants'))
print('cpu', d.get('cpu_baseline', {}).get('value'))
PY
ls -la $O/*.ncu-rep
