#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r50.txt; : > $out
timeout 300 tools/qbench c3 10 "" "pf=888" "pf=1776" "pf=444" "pf=3552" "" "pf=888" "pf=8000" >> $out 2>&1
grep -v "^# " $out
