#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r45.txt; : > $out
for i in 1 2; do
  timeout 300 tools/qbench c3 10 "" >> $out 2>&1
  echo "## pad-per-4 + rotated stores" >> $out
  LD_LIBRARY_PATH=variants/pad4 timeout 300 tools/qbench c3 10 "" >> $out 2>&1
done
timeout 300 tools/qbench w20 5 "" >> $out 2>&1
echo "## pad-per-4 + rotated stores" >> $out
LD_LIBRARY_PATH=variants/pad4 timeout 300 tools/qbench w20 5 "" >> $out 2>&1
grep -v "^# " $out
