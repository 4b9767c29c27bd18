#!/bin/bash
mkdir -p gpurun_out
bash tools/ncu_capture.sh r02_c3 k_wpt 12 4 --workload c3
python tools/ncu_summary.py gpurun_out/prof_r02_c3.raw.csv > gpurun_out/r02_ncu_c3.md
cut -c1-330 gpurun_out/r02_ncu_c3.md
ncu -i /tmp/prof_r02_c3.ncu-rep --page source --csv > gpurun_out/r02_c3.source.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/r02_c3.source.csv 25 > gpurun_out/r02_c3_source_top.txt
bash tools/ncu_capture.sh r02_c2 k_fwt 18 6 --workload c2
python tools/ncu_summary.py gpurun_out/prof_r02_c2.raw.csv > gpurun_out/r02_ncu_c2.md
cut -c1-330 gpurun_out/r02_ncu_c2.md
ncu -i /tmp/prof_r02_c2.ncu-rep --page source --csv > gpurun_out/r02_c2.source.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/r02_c2.source.csv 25 > gpurun_out/r02_c2_source_top.txt
