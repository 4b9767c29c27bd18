#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r42.txt; : > $out
QB_KERNELS=1 timeout 300 tools/qbench c3 10 "" "wpt_tma_store=0" "" "wpt_tma_store=0" >> $out 2>&1
timeout 300 tools/qbench w20 5 "" "wpt_tma_store=0" >> $out 2>&1
timeout 300 tools/qbench w2d 5 "" "wpt_tma_store=0" >> $out 2>&1
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> $out
cat $out
