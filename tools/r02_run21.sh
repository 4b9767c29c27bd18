#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r21.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_inplace=0" >> $out 2>&1
timeout 300 tools/qbench w20 5 "" >> $out 2>&1
QB_BATCH=64 timeout 300 tools/qbench c3 5 "" "wpt_m=2" "wpt_tile=1024,wpt_threads=96" >> $out 2>&1
cat $out
