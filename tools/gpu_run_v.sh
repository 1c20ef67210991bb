#!/bin/bash
# GPU call V: single-pass TF32 with the L2 policies in both chained kernels: shared-memory operand vs TMEM operand per sweep
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout 200 "$@" > $O/v_$name.log 2>&1; echo "$name rc=$?"; }
FBSNN_CHAIN_TA=0 FBSNN_CHAIN_HINT=0 run tf32_smem_hint0 python tools/launch_table.py 65536 tf32
FBSNN_CHAIN_TA=0 run tf32_smem_hint7 python tools/launch_table.py 65536 tf32
FBSNN_CHAIN_TA=2 run tf32_tmem_hint7 python tools/launch_table.py 65536 tf32
FBSNN_CHAIN_TA=0 run tf32_smem_hint7_b python tools/launch_table.py 65536 tf32
FBSNN_CHAIN_TA=2 run tf32_tmem_hint7_b python tools/launch_table.py 65536 tf32
FBSNN_CHAIN_TA=0 run x3_smem_hint7 python tools/launch_table.py 65536 tf32x3
FBSNN_CHAIN_TA=1 run x3_tmem_hint7 python tools/launch_table.py 65536 tf32x3
for f in $O/v_*.log; do echo "== $f"; grep -E "\*|step" $f | head -6; done
