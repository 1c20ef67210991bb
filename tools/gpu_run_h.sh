#!/bin/bash
# GPU call H (round 2): chain kernel (TMA-staged I/O) with the weight k-blocks TMA-multicast over clusters of 2 / 4 CTAs.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/h_summary.txt
run() { name=$1; shift; timeout 300 "$@" > $O/h_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/h_summary.txt; }
export FBSNN_CHAIN_DEBUG=1
run diag_cl2_m3     python tools/chain_diag.py --precision tf32x3 --paths 3 --cluster 2
run diag_cl2_m40    python tools/chain_diag.py --precision tf32x3 --paths 40 --cluster 2
run diag_cl2_m2000  python tools/chain_diag.py --precision tf32x3 --paths 2000 --cluster 2
run diag_cl4_m2000  python tools/chain_diag.py --precision tf32x3 --paths 2000 --cluster 4
run diag_cl2_small  python tools/chain_diag.py --precision tf32x3 --paths 300 --steps 7 --dim 10 --layers 11,64,128,64,1 --act Tanh --cluster 2
run diag_tf32_cl2   python tools/chain_diag.py --precision tf32 --paths 2000 --cluster 2
run diag_tf32_cl4   python tools/chain_diag.py --precision tf32 --paths 2000 --cluster 4
unset FBSNN_CHAIN_DEBUG
for cl in 1 2 4; do
  FBSNN_CHAIN=2 FBSNN_CHAIN_CLUSTER=$cl run table_x3_cl$cl python tools/launch_table.py 65536 tf32x3
  FBSNN_CHAIN=2 FBSNN_CHAIN_CLUSTER=$cl run table_tf32_cl$cl python tools/launch_table.py 65536 tf32
done
FBSNN_CHAIN=2 FBSNN_CHAIN_CLUSTER=2 run table_x3_cl2_m4096 python tools/launch_table.py 4096 tf32x3
FBSNN_CHAIN=2 FBSNN_CHAIN_CLUSTER=2 run table_x3_cl2_m100 python tools/launch_table.py 100 tf32x3
timeout 1200 python -m pytest tests -m gpu -q > $O/h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/h_summary.txt
tail -12 $O/h_pytest.log
cat $O/h_summary.txt
for f in $O/h_diag_*.log; do echo "== $f"; grep -v "^ok" $f | tail -8; done
for f in $O/h_table_*.log; do echo "== $f"; grep -E "\*|step|rror|timed" $f | head -8; done
