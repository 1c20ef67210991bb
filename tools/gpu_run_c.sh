#!/bin/bash
# GPU call C (round 2): CTA-pair chain kernel bring-up, aliasing fix check (tf32), timing tables.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/c_summary.txt
run() { name=$1; shift; timeout 240 "$@" > $O/c_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/c_summary.txt; }
export FBSNN_CHAIN_DEBUG=1
run diag_pair_fwd_m3   python tools/chain_diag.py --precision tf32x3 --paths 3 --fwd-only
run diag_pair_full_m3  python tools/chain_diag.py --precision tf32x3 --paths 3
run diag_pair_full_m40 python tools/chain_diag.py --precision tf32x3 --paths 40
run diag_pair_m2000    python tools/chain_diag.py --precision tf32x3 --paths 2000
run diag_single_m2000  python tools/chain_diag.py --precision tf32x3 --paths 2000 --pair 0
run diag_tf32_m2000    python tools/chain_diag.py --precision tf32 --paths 2000
unset FBSNN_CHAIN_DEBUG
timeout 900 python -m pytest tests/test_chain_gpu.py -m gpu -x -q > $O/c_pytest_chain.log 2>&1; echo "pytest chain rc=$?" | tee -a $O/c_summary.txt
tail -5 $O/c_pytest_chain.log
FBSNN_CHAIN=2 run table_x3_pair   python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN=2 FBSNN_CHAIN_PAIR=0 run table_x3_single python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN=2 run table_tf32_chain python tools/launch_table.py 65536 tf32
FBSNN_CHAIN=2 run table_x3_pair_m100 python tools/launch_table.py 100 tf32x3
FBSNN_CHAIN=2 run table_x3_pair_m4096 python tools/launch_table.py 4096 tf32x3
FBSNN_CHAIN=0 run table_x3_perlayer_m4096 python tools/launch_table.py 4096 tf32x3
cat $O/c_summary.txt
for f in $O/c_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -12; done
for f in $O/c_table_*.log; do echo "== $f"; grep -E "\*|step|G |rror" $f | head -12; done
