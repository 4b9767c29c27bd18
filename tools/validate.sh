#!/bin/bash
# Round validation on one B200: GPU test-suite, smoke, the bench lines of every config and the reference arm.
# Outputs go to gpurun_out/ (scratch); copy what should be kept into profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c4 --batch 4 --no-cpu --no-e2e > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --workload c5 --no-cpu --no-e2e > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("default", "ref", "c3", "c4", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "value", d.get("value"), "fwd", d.get("forward_gsps"), r.get("forward_frac"), "rev", d.get("reverse_gsps"),
              r.get("reverse_frac"), "dominant", r.get("kernel"), r.get("frac"), "e2e", (d.get("e2e") or {}).get("value"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
