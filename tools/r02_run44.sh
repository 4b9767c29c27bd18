#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r44.txt; : > $out
for w in c2 c4 c5 d20; do
  QB_KERNELS=1 timeout 300 tools/qbench $w 5 "" 2>&1 | grep -E "^#|default|k_fwt_rev" >> $out
  echo "## no final stores in k_fwt_rev" >> $out
  QB_KERNELS=1 LD_LIBRARY_PATH=variants/nostg timeout 300 tools/qbench $w 5 "" 2>&1 | grep -E "default|k_fwt_rev" >> $out
done
ncu --set full --import-source on --clock-control none -k regex:k_wpt_rev -s 2 -c 1 -o /tmp/prof_rev tools/qbench c3 1 "" > gpurun_out/r44_ncu.log 2>&1
ncu -i /tmp/prof_rev.ncu-rep --page source --csv > gpurun_out/r44_rev.source.csv 2>/dev/null
cat $out
