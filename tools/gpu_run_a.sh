#!/bin/bash
# GPU call A (round 2): bring-up of the layer-chained sweep kernels.  Every step in its own process with a timeout.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/a_gpu.txt 2>&1
export FBSNN_CHAIN_DEBUG=1
run() { name=$1; shift; timeout 180 "$@" > $O/a_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/a_summary.txt; }
rm -f $O/a_summary.txt
run diag_x3_fwd_m3   python tools/chain_diag.py --precision tf32x3 --paths 3 --fwd-only
run diag_x3_full_m3  python tools/chain_diag.py --precision tf32x3 --paths 3
run diag_x3_full_m40 python tools/chain_diag.py --precision tf32x3 --paths 40
run diag_tf32_full_m40 python tools/chain_diag.py --precision tf32 --paths 40
run diag_x3_small    python tools/chain_diag.py --precision tf32x3 --paths 300 --steps 7 --dim 10 --layers 11,64,128,64,1 --act Tanh
run diag_x3_hjb      python tools/chain_diag.py --precision tf32x3 --paths 77 --steps 12 --dim 20 --layers 21,96,96,1 --act ReLU --problem hjb
unset FBSNN_CHAIN_DEBUG
timeout 900 python -m pytest tests -m gpu -x -q > $O/a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/a_summary.txt
tail -5 $O/a_pytest.log
for prec in tf32x3 tf32; do
  FBSNN_CHAIN=0 timeout 300 python tools/launch_table.py 65536 $prec > $O/a_table_${prec}_perlayer.log 2>&1; echo "table $prec perlayer rc=$?" | tee -a $O/a_summary.txt
  FBSNN_CHAIN=2 timeout 300 python tools/launch_table.py 65536 $prec > $O/a_table_${prec}_chain.log 2>&1; echo "table $prec chain rc=$?" | tee -a $O/a_summary.txt
done
FBSNN_CHAIN=2 timeout 300 python tools/launch_table.py 100 tf32x3 > $O/a_table_x3_m100_chain.log 2>&1
FBSNN_CHAIN=0 timeout 300 python tools/launch_table.py 100 tf32x3 > $O/a_table_x3_m100_perlayer.log 2>&1
cat $O/a_summary.txt
for f in $O/a_diag_*.log; do echo "== $f"; tail -45 $f; done
for f in $O/a_table_*chain.log; do echo "== $f"; tail -12 $f; done
