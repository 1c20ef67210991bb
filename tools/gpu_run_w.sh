#!/bin/bash
# GPU call W (round 2): final state -- full GPU test-suite, smoke, bench (our arm + reference arm), ncu launch list,
# per-launch DRAM traffic of a step, cycle accounting.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/w_summary.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/w_summary.txt
timeout 300 python __graft_entry__.py smoke > $O/w_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/w_summary.txt
for prec in tf32x3 tf32; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/w_traffic_$prec.csv python tools/one_step.py 65536 $prec > $O/w_traffic_$prec.log 2>&1
  echo "traffic $prec rc=$?" | tee -a $O/w_summary.txt
done
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/w_ref.log 2>&1; echo "ref rc=$?" | tee -a $O/w_summary.txt
( time timeout 1500 python bench.py ) > $O/w_bench.log 2>&1; echo "bench rc=$?" | tee -a $O/w_summary.txt
grep '^{' $O/w_bench.log | tail -1 > $O/w_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/w_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-mc --skip-cpu --skip-small --skip-variants --skip-e2e --skip-workloads --skip-peak > $O/w_ncu.log 2>&1
echo "ncu-list rc=$?" | tee -a $O/w_summary.txt
tail -6 $O/w_pytest.log
tail -2 $O/w_smoke.log
cat $O/w_summary.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/w_bench.json'))
for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks', 'e2e', 'e2e_philox', 'e2e_numpy_default', 'small_m', 'mid_m'):
    print(k, str(d.get(k))[:260])
r = d['roofline']
print('roofline', {k: r[k] for k in ('bound', 'achieved', 'peak', 'frac', 'traffic')})
for t in r['launch_table']:
    print(t)
print('variants', {k: v['value'] for k, v in d.get('variants', {}).items()})
for w in d.get('workloads', []):
    print(w.get('kind'), w.get('dim'), w.get('act'), w.get('paths'), w.get('iters_per_s'), w.get('roofline', {}).get('frac'), w.get('error'))
print('mc', d.get('mc', {}).get('value'), d.get('mc', {}).get('roofline', {}).get('frac'))
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'))
PY
