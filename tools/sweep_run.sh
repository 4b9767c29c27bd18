#!/bin/bash
# same-run A/B sweeps on the GPU box: bash tools/sweep_run.sh <log> <workload> <tune> [<tune> ...]  (see tools/sweep.py)
log=$1; shift
python tools/sweep.py "$@" > gpurun_out/$log 2>&1
cat gpurun_out/$log
