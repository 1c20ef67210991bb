#!/bin/bash
# GPU call AA: chaint_kernel store warp with two store groups in flight (FBSNN_CHAIN_HINT bit 3)
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout 200 "$@" > $O/aa_$name.log 2>&1; echo "$name rc=$?"; }
FBSNN_CHAIN_HINT=15 FBSNN_CHAIN_DEBUG=1 run diag15 python tools/chain_diag.py --precision tf32x3 --paths 2000
for h in 7 15 7 15; do FBSNN_CHAIN_HINT=$h run x3_hint${h}_$RANDOM python tools/launch_table.py 65536 tf32x3; done
for h in 7 15; do FBSNN_CHAIN_HINT=$h FBSNN_CHAIN_TA=2 run tf32_hint${h} python tools/launch_table.py 65536 tf32; done
grep -E "DIAG|BAD" $O/aa_diag15.log | head -5
for f in $O/aa_x3_*.log $O/aa_tf32_*.log; do echo "== $f"; grep -E "\*|step" $f | head -6; done
