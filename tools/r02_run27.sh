#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r27.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench d20 5 "" >> $out 2>&1
timeout 300 tools/qbench c4 5 "" >> $out 2>&1
timeout 300 tools/qbench c5 5 "" >> $out 2>&1
timeout 300 tools/qbench c2 5 "" >> $out 2>&1
timeout 300 tools/qbench d10 5 "" >> $out 2>&1
cat $out
