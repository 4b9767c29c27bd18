#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r31.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c5 5 "" "res_split=32" >> $out 2>&1
timeout 300 tools/qbench c5 5 "res_split=16" "res_split=64" "res_split=32,res_threads=256" >> $out 2>&1
timeout 300 tools/qbench d20 5 "" "res_split=32" >> $out 2>&1
timeout 300 tools/qbench c2 5 "" "res_split=32" "res_split=16" >> $out 2>&1
cat $out
