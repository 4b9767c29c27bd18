#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r23.txt; : > $out
timeout 300 tools/qbench c3 10 "" "wpt_rs=4,wpt_r=4,wpt_threads=288" "wpt_rs=4,wpt_r=4,wpt_tile=1024,wpt_threads=160" "wpt_rs=4,wpt_threads=288,wpt_inplace=0" "wpt_rs=4,wpt_r=4,wpt_threads=160,wpt_inplace=0" >> $out 2>&1
cat $out
