#!/bin/bash
# GPU call B (round 2): tf32 chain failure localisation + ablation timings of the chained sweeps.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/b_summary.txt
run() { name=$1; shift; timeout 240 "$@" > $O/b_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/b_summary.txt; }
FBSNN_CHAIN_DEBUG=1 run diag_tf32_m2000 python tools/chain_diag.py --precision tf32 --paths 2000
FBSNN_CHAIN_DEBUG=1 run diag_x3_m2000 python tools/chain_diag.py --precision tf32x3 --paths 2000
for ab in 0 1 2 4 8 16 24 7 3; do
  FBSNN_CHAIN=2 FBSNN_CHAIN_ABLATE=$ab run table_x3_ab$ab python tools/launch_table.py 65536 tf32x3
done
FBSNN_CHAIN=2 run table_tf32_chain python tools/launch_table.py 65536 tf32
cat $O/b_summary.txt
tail -12 $O/b_diag_tf32_m2000.log
for ab in 0 1 2 4 8 16 24 7 3; do echo "== ablate $ab"; grep -E "\*|step" $O/b_table_x3_ab$ab.log; done
tail -12 $O/b_table_tf32_chain.log
