#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r29.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench w2d 5 "" "wpt_transpose=0" >> $out 2>&1
timeout 300 tools/qbench w3d 5 "" "wpt_transpose=0" >> $out 2>&1
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 >> $out
cat $out
