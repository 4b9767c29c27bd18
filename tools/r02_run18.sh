#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r18.txt; : > $out
timeout 120 tools/microbench2 >> $out 2>&1
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_tile=4096,wpt_threads=288" "wpt_tile=1024,wpt_threads=96" >> $out 2>&1
cat $out
