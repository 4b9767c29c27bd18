#!/bin/bash
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r02_bench_all_n1.json 2> gpurun_out/r02_bench_all_n1.err ) 2> gpurun_out/r02_bench_all_n1.time
tail -3 gpurun_out/r02_bench_all_n1.time; tail -5 gpurun_out/r02_bench_all_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_all_n1.json").read().strip().splitlines()[-1])
def show(tag, r):
    if "error" in r: print(tag, "ERROR", r["error"]); return
    rf = r["roofline"]
    print(tag, "value %.1f" % r["value"], "fwd %.3f rev %.3f" % (rf.get("forward_frac", 0), rf.get("reverse_frac", 0)), "dom", rf["kernel"], "%.3f" % rf["frac"],
          "e2e", (r.get("e2e") or {}).get("value"), "pcie frac", (r.get("e2e") or {}).get("frac_of_pcie_ceiling"), "cpu", (r.get("cpu_baseline") or {}).get("value"), "rt", r["roundtrip_max_abs_err"])
show("c2", d)
for k, v in d.get("workloads", {}).items(): show(k, v)
PY
( time python bench.py --impl reference > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err ) 2> gpurun_out/r02_bench_ref_n1.time
tail -3 gpurun_out/r02_bench_ref_n1.time; cut -c1-600 gpurun_out/r02_bench_ref_n1.json
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
