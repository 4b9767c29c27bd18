#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r39.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "dbg=1" "dbg=2" "dbg=4" "dbg=6" "dbg=7" "wpt_rev_m=6,dbg=2" "wpt_rev_m=6,dbg=6" "wpt_rev_m=6,dbg=7" 2>&1 | grep -v "k_wpt_fwd" >> $out
cat $out
