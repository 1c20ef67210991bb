#!/bin/bash
# GPU call G (round 2): chain kernels v3 (register-level global I/O), single-CTA and CTA-pair.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/g_summary.txt
run() { name=$1; shift; timeout 300 "$@" > $O/g_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/g_summary.txt; }
export FBSNN_CHAIN_DEBUG=1
run diag_single_m40   python tools/chain_diag.py --precision tf32x3 --paths 40 --pair 0
run diag_single_m2000 python tools/chain_diag.py --precision tf32x3 --paths 2000 --pair 0
run diag_pair_fwd_m3  python tools/chain_diag.py --precision tf32x3 --paths 3 --fwd-only
run diag_pair_m40     python tools/chain_diag.py --precision tf32x3 --paths 40
run diag_pair_m2000   python tools/chain_diag.py --precision tf32x3 --paths 2000
run diag_tf32_m2000   python tools/chain_diag.py --precision tf32 --paths 2000
run diag_x3_small     python tools/chain_diag.py --precision tf32x3 --paths 300 --steps 7 --dim 10 --layers 11,64,128,64,1 --act Tanh
unset FBSNN_CHAIN_DEBUG
FBSNN_CHAIN=2 run table_x3_pair   python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN=2 FBSNN_CHAIN_PAIR=0 run table_x3_single python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN=2 run table_tf32 python tools/launch_table.py 65536 tf32
FBSNN_CHAIN=2 run table_x3_pair_m100 python tools/launch_table.py 100 tf32x3
FBSNN_CHAIN=2 run table_x3_pair_m4096 python tools/launch_table.py 4096 tf32x3
for ab in 1 2 3 4 8 7; do
  FBSNN_CHAIN=2 FBSNN_CHAIN_ABLATE=$ab run table_x3_pair_ab$ab python tools/launch_table.py 65536 tf32x3
done
timeout 1200 python -m pytest tests -m gpu -q > $O/g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/g_summary.txt
tail -15 $O/g_pytest.log
cat $O/g_summary.txt
for f in $O/g_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -8; done
for f in $O/g_table_*.log; do echo "== $f"; grep -E "\*|step|rror|timed" $f | head -8; done
