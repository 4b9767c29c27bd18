#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r24.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_inplace=0" "wpt_rs=4,wpt_threads=288" >> $out 2>&1
timeout 300 tools/qbench w20 5 "" >> $out 2>&1
QB_BATCH=64 timeout 300 tools/qbench c3 3 "" "wpt_m=2" "wpt_m=1" "wpt_tile=1024,wpt_threads=96" "wpt_tile=512,wpt_threads=64" >> $out 2>&1
cat $out
