#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r25.txt; : > $out
timeout 300 tools/qbench c3 10 "" "stagger=300" "stagger=600" "stagger=1000" "stagger=1500" "stagger=2500" "stagger=5000" >> $out 2>&1
timeout 300 tools/qbench w20 5 "" >> $out 2>&1
cat $out
