#!/bin/bash
# N GPUs: c5 slab curve variants
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1; free -g | head -2 >> gpurun_out/r02_topo_n$N.txt; nproc >> gpurun_out/r02_topo_n$N.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q --tb=short 2>&1 | tail -5
run() { # tag, env..., extra args
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 $EXTRA > gpurun_out/r02_bench_c5_n${N}_$tag.json 2> gpurun_out/r02_bench_c5_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c5_n${N}_$tag.json").read().strip().splitlines()[-1])
    s = d.get("slab", {})
    print("$tag: value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in s.items() if k not in ("note", "mode")}, d.get("slab_parity_max_err"), d.get("slab_parity", {}).get("full_tol"))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02_bench_c5_n${N}_$tag.err").read()[-1500:])
PY
}
EXTRA="" run j_ch4 JWB_SLAB_CHUNKS=4
EXTRA="" run j_ch8 JWB_SLAB_CHUNKS=8
EXTRA="" run j_ch2 JWB_SLAB_CHUNKS=2
EXTRA="--slab-layout i" run i_ch4 JWB_SLAB_CHUNKS=4
EXTRA="--slab all2all" run a2a JWB_SLAB_CHUNKS=4
