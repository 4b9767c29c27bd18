#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "strided or 2d or 3d" 2>&1 | tail -3
SWEEP_BATCH=1 timeout 600 python tools/sweep.py c5 "" "str2_cap=256" "str2_cap=128" > gpurun_out/r02_run14_sweep_c5.txt 2>&1
cat gpurun_out/r02_run14_sweep_c5.txt
SWEEP_BATCH=16 timeout 600 python tools/sweep.py c4 "" "str2_cap=256" > gpurun_out/r02_run14_sweep_c4.txt 2>&1
cat gpurun_out/r02_run14_sweep_c4.txt
python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_run14_c5.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_run14_c5.json').read().strip().splitlines()[-1]); r=d["roofline"]
for k in r["kernels"]: print("   ",k["kernel"],k["launches"],"avg",round(k["avg_ms"],3),"share",round(k["share"],3),"lv",k["levels"],"fp64",round(k["fp64_frac"],3))
PY
