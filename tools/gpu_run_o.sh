#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
FBSNN_LIB_PATH=build/libfbsnn_prof.so timeout 200 python tools/chain_diag.py --precision tf32x3 --paths 2000 > $O/o_dbg_m2000.log 2>&1
grep -v "^ok" $O/o_dbg_m2000.log | head -70
echo ======== PDL off
FBSNN_PDL=0 FBSNN_LIB_PATH=build/libfbsnn_prof.so timeout 200 python tools/chain_diag.py --precision tf32x3 --paths 2000 > $O/o_dbg_m2000_nopdl.log 2>&1
grep -v "^ok" $O/o_dbg_m2000_nopdl.log | grep -v "^     " | head -30
echo ======== fwd only 2000
FBSNN_LIB_PATH=build/libfbsnn_prof.so timeout 200 python tools/chain_diag.py --precision tf32x3 --paths 2000 --fwd-only > $O/o_dbg_m2000_fwd.log 2>&1
grep -v "^ok" $O/o_dbg_m2000_fwd.log | head -30
