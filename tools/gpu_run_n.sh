#!/bin/bash
# GPU call N (round 2): specialised half-team chaint_kernel -- timeout record of the T sweep, launch tables, cycle profile.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/n_summary.txt
run() { name=$1; shift; timeout 200 "$@" > $O/n_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/n_summary.txt; }
FBSNN_LIB_PATH=build/libfbsnn_prof.so FBSNN_CHAIN_DEBUG=1 run dbg_x3_m2000 python tools/chain_diag.py --precision tf32x3 --paths 2000
FBSNN_LIB_PATH=build/libfbsnn_prof.so FBSNN_CHAIN_DEBUG=1 run dbg_x3_m700 python tools/chain_diag.py --precision tf32x3 --paths 700
FBSNN_CHAIN_TA=1 run table_x3_ta1 python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN_TA=1 run table_tf32_ta1 python tools/launch_table.py 65536 tf32
FBSNN_CHAIN_TA=0 run table_x3_ta0 python tools/launch_table.py 65536 tf32x3
run prof_x3 python tools/chain_prof.py 65536 tf32x3
cat $O/n_summary.txt
for f in $O/n_dbg_*.log; do echo "== $f"; grep -v "^ok" $f | tail -12; done
for f in $O/n_table_*.log; do echo "== $f"; grep -E "\*|step|rror|timed" $f | head -8; done
grep "mma:" $O/n_prof_x3.log
