#!/bin/bash
# per-kernel tables of every workload with the final kernels (pairs with profiles/r02_qbench_start.txt)
mkdir -p gpurun_out; out=gpurun_out/r51.txt; : > $out
export QB_KERNELS=1
for w in c2 c3 c4 c5 d20 w20 h1; do timeout 120 tools/qbench $w 10 "" >> $out 2>&1; done
cat $out
