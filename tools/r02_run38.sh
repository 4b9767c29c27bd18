#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r38.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_rev_m=6" "wpt_rev_m=4" "wpt_rev_m=5" "wpt_rev_m=6,wpt_tile=4096,wpt_threads=288" 2>&1 | grep -v "k_wpt_fwd" >> $out
cat $out
