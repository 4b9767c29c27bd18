#!/bin/bash
# round 2, GPU run 1: parity of the second-generation strided kernels + A/B against the first generation
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_env.txt; nproc >> gpurun_out/r02_env.txt; free -g >> gpurun_out/r02_env.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "strided or 2d or 3d or axis" 2>&1 | tail -15 > gpurun_out/r02_run1_pytest.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -k "config4 or config5" 2>&1 | tail -8 >> gpurun_out/r02_run1_pytest.log
cat gpurun_out/r02_run1_pytest.log
timeout 900 python tools/sweep.py c4 "" "str_v2=0" "str2_tile=256,str2_rev_tile=256" "str2_cap=256" > gpurun_out/r02_run1_sweep_c4.txt 2>&1
cat gpurun_out/r02_run1_sweep_c4.txt
timeout 900 python tools/sweep.py c5 "" "str_v2=0" "str2_m=3" "str2_cap=256" "str2_m=3,str2_rev_m=3" > gpurun_out/r02_run1_sweep_c5.txt 2>&1
cat gpurun_out/r02_run1_sweep_c5.txt
python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_run1_c4.json 2>gpurun_out/r02_run1_c4.err
python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_run1_c5.json 2>gpurun_out/r02_run1_c5.err
