#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:k_wpt -s 4 -c 4 -o /tmp/prof_c3q tools/qbench c3 1 "" > gpurun_out/r22_ncu.log 2>&1
ncu -i /tmp/prof_c3q.ncu-rep --page raw --csv > gpurun_out/r22_c3.raw.csv 2>/dev/null
ncu -i /tmp/prof_c3q.ncu-rep --page source --csv > gpurun_out/r22_c3.source.csv 2>/dev/null
ls -la gpurun_out/r22*; tail -3 gpurun_out/r22_ncu.log
