#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r43.txt; : > $out
timeout 300 tools/qbench c3 30 "" "wpt_tma_store_fwd=1" "" "wpt_tma_store_fwd=1" "" "wpt_tma_store_fwd=1" >> $out 2>&1
cat $out
