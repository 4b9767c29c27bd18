#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r36.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_warp=1,wpt_tile=512" "wpt_tile=512,wpt_threads=64" 2>&1 | grep -v "k_wpt_rev" >> $out
cat $out
