#!/bin/bash
# GPU call D (round 2): CTA-pair chain kernel (cta_group::2 TMA), full GPU test-suite, new bench.py trial.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/d_summary.txt
run() { name=$1; shift; timeout 300 "$@" > $O/d_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/d_summary.txt; }
export FBSNN_CHAIN_DEBUG=1
run diag_pair_fwd_m3   python tools/chain_diag.py --precision tf32x3 --paths 3 --fwd-only
run diag_pair_full_m40 python tools/chain_diag.py --precision tf32x3 --paths 40
run diag_pair_m2000    python tools/chain_diag.py --precision tf32x3 --paths 2000
unset FBSNN_CHAIN_DEBUG
FBSNN_CHAIN=2 run table_x3_pair   python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN=2 run table_x3_pair_m100 python tools/launch_table.py 100 tf32x3
FBSNN_CHAIN=2 run table_x3_pair_m4096 python tools/launch_table.py 4096 tf32x3
timeout 900 python -m pytest tests -m gpu -q -x > $O/d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/d_summary.txt
tail -8 $O/d_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --mc-paths 268435456 > $O/d_bench.json 2> $O/d_bench.err; echo "bench rc=$?" | tee -a $O/d_summary.txt
tail -3 $O/d_bench.err
cat $O/d_summary.txt
for f in $O/d_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -8; done
for f in $O/d_table_*.log; do echo "== $f"; grep -E "\*|step|G |rror|timed" $f | head -12; done
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/d_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
    print('roofline', {k:d['roofline'][k] for k in ('achieved','peak','frac')})
    for r in d['roofline']['launch_table']: print(r)
    for k in ('e2e','e2e_philox','e2e_numpy_default','small_m','mid_m'): print(k, d.get(k))
    print('variants', d.get('variants'))
    for w in d.get('workloads',[]): print(w)
    print('mc', d.get('mc'))
except Exception as e: print('bench parse failed', e)
PY
