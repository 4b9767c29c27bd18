#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r48.txt; : > $out
timeout 300 tools/qbench d20 5 "" "rev_tile=2048" "rev_tile=2048,rev_m=2" "rev_tile=1024,rev_m=2" "rev_tile=8192" "rev_tile=2048,rev_threads=64" "rev_tile=2048,rev_rs=8" "rev_rs=8" >> $out 2>&1
timeout 300 tools/qbench c5 5 "" "rev_tile=512,rev_m=2" "rev_tile=512,rev_m=2,rev_threads=64" "rev_threads=64" "rev_rs=8,rev_threads=64" >> $out 2>&1
grep -v "^# " $out
