#!/bin/bash
# baseline of every workload through the native A/B driver + shuffle A/B + WPT register-budget variants
mkdir -p gpurun_out; out=gpurun_out/r16.txt; : > $out
export QB_KERNELS=1
for w in c2 c3 c4 c5 d20; do timeout 300 tools/qbench $w 10 "" >> $out 2>&1; done
unset QB_KERNELS
timeout 300 tools/qbench h1 10 "" "shfl=0" "" "shfl=0" >> $out 2>&1
for v in mb7 mb8; do echo "## variant $v" >> $out; LD_LIBRARY_PATH=variants/$v timeout 300 tools/qbench c3 10 "" >> $out 2>&1; done
echo "## regular again" >> $out; timeout 300 tools/qbench c3 10 "" "wpt_inplace=0" >> $out 2>&1
cat $out
