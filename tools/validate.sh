#!/bin/bash
# Round validation on one B200: GPU test-suite, smoke, the default bench line (all four configs), the reference arm,
# launch lists of every config and one ncu --set full capture per config.  Outputs go to gpurun_out/ (scratch);
# copy what should be kept into profiles/.
#   bash tools/validate.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,driver_version --format=csv > gpurun_out/${tag}_env.txt
nproc >> gpurun_out/${tag}_env.txt; free -g | head -2 >> gpurun_out/${tag}_env.txt; (java -version 2>&1 | head -1) >> gpurun_out/${tag}_env.txt
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest_gpu.txt
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.txt 2>&1
( time python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err ) 2> gpurun_out/${tag}_bench_n1.time
( time python bench.py --impl reference > gpurun_out/${tag}_bench_ref_n1.json 2> gpurun_out/${tag}_bench_ref_n1.err ) 2> gpurun_out/${tag}_bench_ref_n1.time
tail -n 3 gpurun_out/${tag}_pytest_gpu.txt gpurun_out/${tag}_smoke.txt gpurun_out/${tag}_bench_n1.time gpurun_out/${tag}_bench_ref_n1.time
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
def show(tag, r):
    if "error" in r: print(tag, "ERROR", r["error"]); return
    rf = r["roofline"]
    print(tag, "value %.1f" % r["value"], "fwd %.3f rev %.3f" % (rf.get("forward_frac", 0), rf.get("reverse_frac", 0)), "dom", rf["kernel"], "%.3f" % rf["frac"],
          "e2e %.2f" % (r.get("e2e") or {}).get("value", 0), "pcie frac %.2f" % (r.get("e2e") or {}).get("frac_of_pcie_ceiling", 0), "cpu %.3f" % (r.get("cpu_baseline") or {}).get("value", 0), "rt", r["roundtrip_max_abs_err"], r.get("clocks"))
show("c2", d)
for k, v in d.get("workloads", {}).items(): show(k, v)
PY
# launch lists (after the plain runs above exited 0) and one full capture per config
for w in c2 c3; do bash tools/ncu_launches.sh ${tag}_$w --workload $w > /dev/null; done
bash tools/ncu_launches.sh ${tag}_c4 --workload c4 --batch 4 > /dev/null
bash tools/ncu_launches.sh ${tag}_c5 --workload c5 > /dev/null
bash tools/ncu_capture.sh ${tag}_c4 k_fwt 39 13 --workload c4 --batch 4 > /dev/null
python tools/ncu_summary.py gpurun_out/prof_${tag}_c4.raw.csv --json gpurun_out/${tag}_traffic_c4.json > gpurun_out/${tag}_ncu_c4.md
bash tools/ncu_capture.sh ${tag}_c5 k_fwt 24 8 --workload c5 > /dev/null
python tools/ncu_summary.py gpurun_out/prof_${tag}_c5.raw.csv --json gpurun_out/${tag}_traffic_c5.json > gpurun_out/${tag}_ncu_c5.md
bash tools/ncu_capture.sh ${tag}_c3 k_wpt 12 4 --workload c3 > /dev/null
python tools/ncu_summary.py gpurun_out/prof_${tag}_c3.raw.csv --json gpurun_out/${tag}_traffic_c3.json > gpurun_out/${tag}_ncu_c3.md
bash tools/ncu_capture.sh ${tag}_c2 k_fwt 18 6 --workload c2 > /dev/null
python tools/ncu_summary.py gpurun_out/prof_${tag}_c2.raw.csv --json gpurun_out/${tag}_traffic_c2.json > gpurun_out/${tag}_ncu_c2.md
rm -f gpurun_out/prof_${tag}_*.raw.csv
cut -c1-250 gpurun_out/${tag}_ncu_c5.md
