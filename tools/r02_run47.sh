#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r47.txt; : > $out
for w in d20 c5; do
  for v in "" nold nofma nofmast; do
    echo "## variant '$v'" >> $out
    QB_KERNELS=1 LD_LIBRARY_PATH=variants/$v timeout 300 tools/qbench $w 5 "" 2>&1 | grep -E "^#|k_fwt_rev:" >> $out
  done
done
cat $out
