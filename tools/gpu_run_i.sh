#!/bin/bash
# GPU call I (round 2): ncu --set full of the chained sweep kernels (one training step, M = 65 536, 3xTF32).
mkdir -p gpurun_out
O=gpurun_out
export FBSNN_CHAIN=2 FBSNN_CHAIN_CLUSTER=1
timeout 300 python tools/one_step.py 65536 tf32x3 > $O/i_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 4 -c 4 -f -o $O/r02_chain_full python tools/one_step.py 65536 tf32x3 > $O/i_ncu.log 2>&1
echo "ncu rc=$?"
tail -5 $O/i_plain.log; tail -5 $O/i_ncu.log; ls -la $O/*.ncu-rep
