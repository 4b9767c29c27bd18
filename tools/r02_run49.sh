#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r49.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_pers=5" "wpt_pers=4" "wpt_pers=6" "wpt_pers=3" "" "wpt_pers=5" 2>&1 | grep -v "k_wpt_rev" >> $out
grep -v "^# " $out
