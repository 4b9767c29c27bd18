#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "1d_parity or batch_parity or ref_" 2>&1 | tail -3
timeout 600 python tools/sweep.py c3 "" "rot_warps=0" "" "rot_warps=0" > gpurun_out/r02_run10_sweep_c3.txt 2>&1
cat gpurun_out/r02_run10_sweep_c3.txt
timeout 600 python tools/sweep.py c2 "" "rot_warps=0" "" "rot_warps=0" > gpurun_out/r02_run10_sweep_c2.txt 2>&1
cat gpurun_out/r02_run10_sweep_c2.txt
