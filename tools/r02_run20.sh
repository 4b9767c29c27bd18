#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r20.txt; : > $out
# 20.6 KB (fwd) / 22.9 KB (rev) per CTA: xsmem k -> CTAs per SM = floor(227 / (21..23 + k))
timeout 300 tools/qbench c3 10 "" "carve=1,xsmem=16" "carve=1,xsmem=22" "carve=1,xsmem=32" "carve=1,xsmem=52" "carve=1,xsmem=90" "xsmem=8" "xsmem=16" >> $out 2>&1
cat $out
