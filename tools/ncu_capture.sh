#!/bin/bash
# One ncu --set full capture of the kernels of ONE bench step (run the plain command first, as the
# profiling recipe asks); the report stays on the box, the raw CSV page comes back in gpurun_out/.
#   bash tools/ncu_capture.sh <tag> <kernel-regex> <launches-to-skip> <launches-to-capture> <bench args...>
tag=$1; regex=$2; skip=$3; count=$4; shift 4
python bench.py "$@" --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:$regex -s $skip -c $count -o /tmp/prof_$tag \
    python bench.py "$@" --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_$tag.raw.csv 2>/dev/null
ls -la gpurun_out/prof_$tag.raw.csv
