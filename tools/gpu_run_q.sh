#!/bin/bash
# GPU call Q: DRAM traffic of every kernel of one training step (M = 65 536, 3xTF32 and TF32), three ncu metrics only.
mkdir -p gpurun_out
O=gpurun_out
for prec in tf32x3 tf32; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/q_traffic_$prec.csv python tools/one_step.py 65536 $prec > $O/q_traffic_$prec.log 2>&1
  echo "traffic $prec rc=$?"
done
wc -l $O/q_traffic_*.csv
