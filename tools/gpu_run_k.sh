#!/bin/bash
# GPU call K (round 2): chaint_kernel (operand of the next MMA in tensor memory) -- bring-up diagnostics, launch tables
# against the shared-memory chain kernel on the same box, ablations.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/k_summary.txt
run() { name=$1; shift; timeout 300 "$@" > $O/k_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/k_summary.txt; }
export FBSNN_CHAIN_DEBUG=1
run diag_x3_fwd_m3  python tools/chain_diag.py --precision tf32x3 --paths 3 --fwd-only
run diag_x3_m3      python tools/chain_diag.py --precision tf32x3 --paths 3
run diag_x3_m40     python tools/chain_diag.py --precision tf32x3 --paths 40
run diag_x3_m2000   python tools/chain_diag.py --precision tf32x3 --paths 2000
run diag_x3_small   python tools/chain_diag.py --precision tf32x3 --paths 300 --steps 7 --dim 10 --layers 11,64,128,64,1 --act Tanh
run diag_tf32_m40   python tools/chain_diag.py --precision tf32 --paths 40
run diag_tf32_m2000 python tools/chain_diag.py --precision tf32 --paths 2000
unset FBSNN_CHAIN_DEBUG
if grep -q "rc=[^0]" $O/k_summary.txt; then
  for f in $O/k_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -12; done
  exit 1
fi
for ta in 1 0 1; do
  FBSNN_CHAIN_TA=$ta run table_x3_ta$ta python tools/launch_table.py 65536 tf32x3
  FBSNN_CHAIN_TA=$ta run table_tf32_ta$ta python tools/launch_table.py 65536 tf32
done
for ab in 1 2 3 4 8 7; do
  FBSNN_CHAIN_ABLATE=$ab run table_x3_ablate$ab python tools/launch_table.py 65536 tf32x3
done
cat $O/k_summary.txt
for f in $O/k_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -4; done
for f in $O/k_table_*.log; do echo "== $f"; grep -E "\*|step|rror|timed" $f | head -8; done
