#!/bin/bash
# GPU call R: chaint_kernel L2 prefetch distance and eviction hints (launch tables, 3xTF32, M = 65 536) + small-M chained dispatch
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout 200 "$@" > $O/r_$name.log 2>&1; echo "$name rc=$?"; }
for pf in 7 0 2 4; do FBSNN_CHAIN_PF=$pf run pf$pf python tools/launch_table.py 65536 tf32x3; done
for h in 1 2 3; do FBSNN_CHAIN_HINT=$h run hint$h python tools/launch_table.py 65536 tf32x3; done
FBSNN_CHAIN_PF=3 FBSNN_CHAIN_HINT=1 run pf3_hint1 python tools/launch_table.py 65536 tf32x3
run m100_default python tools/launch_table.py 100 tf32x3
FBSNN_CHAIN=2 run m100_chain python tools/launch_table.py 100 tf32x3
run m4096_default python tools/launch_table.py 4096 tf32x3
for f in $O/r_*.log; do echo "== $f"; grep -E "\*|step|rror|timed|dense total" $f | head -8; done
