#!/bin/bash
# GPU call M (round 2): chaint_kernel with two half-teams -- diagnostics, launch tables (TA on / off on the same box), cycle profile.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/m_summary.txt
run() { name=$1; shift; timeout 300 "$@" > $O/m_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/m_summary.txt; }
export FBSNN_CHAIN_DEBUG=1
run diag_x3_m3      python tools/chain_diag.py --precision tf32x3 --paths 3
run diag_x3_m40     python tools/chain_diag.py --precision tf32x3 --paths 40
run diag_x3_m2000   python tools/chain_diag.py --precision tf32x3 --paths 2000
run diag_x3_small   python tools/chain_diag.py --precision tf32x3 --paths 300 --steps 7 --dim 10 --layers 11,64,128,64,1 --act Tanh
run diag_x3_odd     python tools/chain_diag.py --precision tf32x3 --paths 77 --steps 12 --dim 20 --layers 21,96,96,1 --act ReLU --problem hjb
run diag_tf32_m2000 python tools/chain_diag.py --precision tf32 --paths 2000
unset FBSNN_CHAIN_DEBUG
if grep -q "rc=[^0]" $O/m_summary.txt; then
  for f in $O/m_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -12; done
  exit 1
fi
for ta in 1 0 1; do
  FBSNN_CHAIN_TA=$ta run table_x3_ta$ta python tools/launch_table.py 65536 tf32x3
  FBSNN_CHAIN_TA=$ta run table_tf32_ta$ta python tools/launch_table.py 65536 tf32
done
run prof_x3 python tools/chain_prof.py 65536 tf32x3
run prof_tf32 python tools/chain_prof.py 65536 tf32
cat $O/m_summary.txt
for f in $O/m_table_*.log; do echo "== $f"; grep -E "\*|step|rror|timed" $f | head -8; done
cat $O/m_prof_x3.log
