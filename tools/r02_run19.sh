#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r19.txt; : > $out
timeout 300 tools/qbench c3 10 "" "carve=1" >> $out 2>&1
for v in mb7 mb8; do echo "## variant $v" >> $out; LD_LIBRARY_PATH=variants/$v timeout 300 tools/qbench c3 10 "" "carve=1" >> $out 2>&1; done
timeout 300 tools/qbench c4 10 "" "carve=1" >> $out 2>&1
timeout 300 tools/qbench c5 10 "" "carve=1" >> $out 2>&1
cat $out
