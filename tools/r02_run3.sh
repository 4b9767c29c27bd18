#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "strided" 2>&1 | tail -3
timeout 600 python tools/sweep.py c4 "" "str2_tile=256" "str2_rev_tile=256" "str2_tile=256,str2_rev_tile=256" > gpurun_out/r02_run3_sweep_c4.txt 2>&1
cat gpurun_out/r02_run3_sweep_c4.txt
timeout 600 python tools/sweep.py c5 "" "str2_tile=256" "str2_rev_tile=256" "str2_m=3" > gpurun_out/r02_run3_sweep_c5.txt 2>&1
cat gpurun_out/r02_run3_sweep_c5.txt
bash tools/ncu_capture.sh r02_c4_str2b str2 18 6 --workload c4 --batch 4
python tools/ncu_summary.py gpurun_out/prof_r02_c4_str2b.raw.csv > gpurun_out/r02_ncu_c4_str2b.md
cut -c1-330 gpurun_out/r02_ncu_c4_str2b.md
