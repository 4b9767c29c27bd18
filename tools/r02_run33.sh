#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r33.txt; : > $out
show='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
def one(t,r):
    rf=r["roofline"]; print(t, "fwd %.3f rev %.3f"%(rf["forward_frac"],rf["reverse_frac"]), r.get("clocks"))
one("c2",d)
for k,v in d.get("workloads",{}).items(): one(k,v)
'
echo "== all, no cpu / e2e" >> $out; python bench.py --no-cpu --no-e2e 2>/dev/null | python -c "$show" >> $out 2>&1
echo "== c3 alone" >> $out; python bench.py --workload c3 --no-cpu --no-e2e 2>/dev/null | python -c "$show" >> $out 2>&1
echo "== all, with cpu + e2e" >> $out; python bench.py 2>/dev/null | python -c "$show" >> $out 2>&1
cat $out
