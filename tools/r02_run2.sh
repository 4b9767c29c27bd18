#!/bin/bash
# round 2, GPU run 2: reverse issue-loop fix A/B + ncu of the strided2 kernels (C4, 4 images)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "strided" 2>&1 | tail -3
timeout 600 python tools/sweep.py c4 "" "str2_tile=256,str2_rev_tile=256" "str2_rev_tile=256" > gpurun_out/r02_run2_sweep_c4.txt 2>&1
cat gpurun_out/r02_run2_sweep_c4.txt
timeout 600 python tools/sweep.py c5 "" "str2_tile=256,str2_rev_tile=256" > gpurun_out/r02_run2_sweep_c5.txt 2>&1
cat gpurun_out/r02_run2_sweep_c5.txt
# ncu: one step of c4 with 4 images; kernels named str2 (fwd tile x2, fwd res, rev res, rev tile x2 = 6 per step); skip 3 warm-up steps
bash tools/ncu_capture.sh r02_c4_str2 str2 18 6 --workload c4 --batch 4
python tools/ncu_summary.py gpurun_out/prof_r02_c4_str2.raw.csv > gpurun_out/r02_ncu_c4_str2.md
cut -c1-330 gpurun_out/r02_ncu_c4_str2.md
ncu -i /tmp/prof_r02_c4_str2.ncu-rep --page source --csv -k regex:k_fwt_rev_str2 > gpurun_out/r02_c4_str2_rev.source.csv 2>/dev/null
ls -la gpurun_out/*.source.csv | tail -2
rm -f gpurun_out/prof_r02_c4_str2.raw.csv.tmp
