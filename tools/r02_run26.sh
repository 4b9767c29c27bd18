#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r26.txt; : > $out
timeout 300 tools/qbench c5 5 "" "res_kb=32" "res_kb=24" "res_kb=16" "res_kb=12" "res_kb=24,res_threads=64" "res_kb=24,res_threads=256" >> $out 2>&1
timeout 300 tools/qbench c4 5 "" "res_kb=24" "res_kb=16" >> $out 2>&1
timeout 300 tools/qbench c2 5 "" "res_kb=24" "res_kb=16" >> $out 2>&1
timeout 300 tools/qbench c3 5 "" >> $out 2>&1
cat $out
