#!/bin/bash
# Launch list (gpu__time_duration only) of one bench command, after the plain run exited 0.
#   bash tools/ncu_launches.sh <tag> <bench args...>
tag=$1; shift
python bench.py "$@" --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py "$@" --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$tag.log 2>&1
tail -n 12 gpurun_out/launches_$tag.csv | cut -c1-200
