#!/bin/bash
# GPU call F (round 2): CTA-pair chain kernel v2 (decoupled input ring), full test-suite, bench trial.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/f_summary.txt
run() { name=$1; shift; timeout 300 "$@" > $O/f_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/f_summary.txt; }
export FBSNN_CHAIN_DEBUG=1
run diag_pair_fwd_m3  python tools/chain_diag.py --precision tf32x3 --paths 3 --fwd-only
run diag_pair_m40     python tools/chain_diag.py --precision tf32x3 --paths 40
run diag_pair_m2000   python tools/chain_diag.py --precision tf32x3 --paths 2000
unset FBSNN_CHAIN_DEBUG
FBSNN_CHAIN=2 run table_x3_pair   python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN=2 run table_x3_pair_m100 python tools/launch_table.py 100 tf32x3
FBSNN_CHAIN=2 run table_x3_pair_m4096 python tools/launch_table.py 4096 tf32x3
timeout 1200 python -m pytest tests -m gpu -q > $O/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/f_summary.txt
tail -15 $O/f_pytest.log
cat $O/f_summary.txt
for f in $O/f_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -8; done
for f in $O/f_table_*.log; do echo "== $f"; grep -E "\*|step|rror|timed|G " $f | head -10; done
