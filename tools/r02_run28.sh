#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r28.txt; : > $out
timeout 300 tools/qbench c5 5 "" "res_kb=64" "res_kb=96" "res_kb=96,res_threads=256" "res_cap=128" "res_cap=64" "res_cap=512,res_kb=96,res_threads=256" >> $out 2>&1
cat $out
