#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r41.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_tma_store=0" "" "wpt_tma_store=0" "wpt_rev_m=6" "wpt_rev_m=6,wpt_tma_store=0" 2>&1 | grep -v "k_wpt_fwd" >> $out
timeout 300 tools/qbench w20 5 "" "wpt_tma_store=0" >> $out 2>&1
cat $out
