#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r37.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "dbg=1" "dbg=2" "dbg=3" 2>&1 | grep -v "k_wpt_rev" >> $out
cat $out
