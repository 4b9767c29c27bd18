#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r40.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "dbg=7" "dbg=15" "dbg=8" "dbg=9" "dbg=14" 2>&1 | grep -v "k_wpt_fwd" >> $out
cat $out
