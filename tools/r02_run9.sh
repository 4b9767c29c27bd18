#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q --tb=short 2>&1 | tail -15
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 $EXTRA > gpurun_out/r02_bench_c5_n${N}_$tag.json 2> gpurun_out/r02_bench_c5_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c5_n${N}_$tag.json").read().strip().splitlines()[-1])
    s = d.get("slab", {})
    print("$tag: value %.1f ms/step %.2f fwd %.2f rev %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["forward_ms"], d["roofline"]["reverse_ms"]), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in s.items() if k not in ("note", "mode", "bytes_sent_per_gpu_per_exchange", "exchanges_per_direction")}, d.get("slab_parity_max_err"))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02_bench_c5_n${N}_$tag.err").read()[-1500:])
PY
}
EXTRA=""
run t_ch4 JWB_SLAB_CHUNKS=4
run t_ch2 JWB_SLAB_CHUNKS=2
run t_ch8 JWB_SLAB_CHUNKS=8
