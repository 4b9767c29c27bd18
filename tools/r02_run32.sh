#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r32.txt; : > $out
timeout 300 tools/qbench c3 20 "" >> $out 2>&1
for per in 0.005 0.05 1000; do
  JWB_CLOCK_PERIOD=$per python bench.py --workload c3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('bench c3 period $per', 'fwd %.3f rev %.3f'%(r['forward_frac'],r['reverse_frac']), [(k['kernel'],round(k['fp64_frac'],3)) for k in r['kernels']], d['clocks'])" >> $out 2>&1
done
timeout 300 tools/qbench c3 20 "" >> $out 2>&1
cat $out
