#!/bin/bash
# GPU call S: weights with L2 evict_last in chaint_kernel: launch tables + DRAM traffic
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout 200 "$@" > $O/s_$name.log 2>&1; echo "$name rc=$?"; }
for h in 0 4 5 7 0 4; do FBSNN_CHAIN_HINT=$h run hint${h}_$RANDOM python tools/launch_table.py 65536 tf32x3; done
for h in 0 4; do
FBSNN_CHAIN_HINT=$h timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv -k regex:chaint_kernel \
    --log-file $O/s_traffic_hint$h.csv python tools/one_step.py 16384 tf32x3 > $O/s_traffic_hint$h.log 2>&1
done
for f in $O/s_hint*.log; do echo "== $f"; grep -E "\*|step" $f | head -6; done
grep -h "chaint" $O/s_traffic_hint0.csv | cut -d, -f5,13- | tail -12
echo ---
grep -h "chaint" $O/s_traffic_hint4.csv | cut -d, -f5,13- | tail -12
