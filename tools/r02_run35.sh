#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r35.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_r=16,wpt_threads=96" "wpt_r=16,wpt_tile=4096,wpt_threads=160" "wpt_r=16,wpt_tile=1024,wpt_threads=64" 2>&1 | grep -v "k_wpt_rev" >> $out
timeout 300 tools/qbench c2 10 "" "res_cap=512" "res_cap=1024" "res_cap=2048" "res_cap=1024,rev_tile=2048" >> $out 2>&1
cat $out
