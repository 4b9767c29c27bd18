#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n${N}_group.json 2> gpurun_out/r02_bench_c5_n${N}_group.err
tail -3 gpurun_out/r02_bench_c5_n${N}_group.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_c5_n${N}_group.json").read().strip().splitlines()[-1])
print("value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), "e2e:", json.dumps(d.get("e2e")))
PY
