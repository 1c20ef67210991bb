#!/bin/bash
# GPU call U: where do the extra DRAM reads of the A / T sweeps come from?  ncu DRAM counters with stores / loads ablated.
mkdir -p gpurun_out
O=gpurun_out
for ab in 0 1 2 16; do
FBSNN_CHAIN_ABLATE=$ab timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none --csv -k regex:chaint_kernel \
    --log-file $O/u_ablate$ab.csv python tools/one_step.py 16384 tf32x3 > $O/u_ablate$ab.log 2>&1
echo "== ablate $ab"
grep -h "chaint" $O/u_ablate$ab.csv | tail -20 | awk -F'","' '{print $5, $(NF-2), $NF}' | tr -d '"'
done
