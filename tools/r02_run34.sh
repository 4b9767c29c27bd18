#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r34.txt; : > $out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "transposes or shuffle or egyptian" 2>&1 | tail -3 >> $out
QB_KERNELS=1 timeout 300 tools/qbench c5 5 "" "res_cap=1024,res_split=256" "res_cap=1024,res_split=256,res_kb=64" "res_cap=1024,res_split=256,res_kb=32" "res_cap=1024,res_split=256,res_threads=256,res_kb=96" 2>&1 | grep -v "str2\|k_fwt_rev" >> $out
cat $out
