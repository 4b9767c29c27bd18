#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r46.txt; : > $out
timeout 300 tools/qbench c2 10 "" "rev_m=4" "rev_m=4,rev_tile=8192" "rev_m=4,rev_threads=160" "rev_m=2" "fwd_m=5" "fwd_tile=4096,fwd_m=5" "rev_m=4,res_cap=512" >> $out 2>&1
grep -v "^# " $out
