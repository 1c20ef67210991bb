#!/bin/bash
# GPU call AD: L2 policies at small / mid batch sizes (row arrays fit in L2 there)
mkdir -p gpurun_out
O=gpurun_out
for M in 100 1000 4096; do for h in 7 4 0 7 4; do
  FBSNN_CHAIN_HINT=$h timeout 200 python tools/launch_table.py $M tf32x3 > $O/ad_m${M}_h${h}_$RANDOM.log 2>&1
done; done
for f in $O/ad_*.log; do echo "== $f $(grep -E 'dense total' $f) $(grep -E '^M=' $f | cut -d: -f2)"; done
