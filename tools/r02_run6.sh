#!/bin/bash
# 2 GPUs: slab-decomposed c5 (chunked copies), multi-GPU tests, and the full default line at N = 2
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1; free -g | head -2 >> gpurun_out/r02_topo_n$N.txt; nproc >> gpurun_out/r02_topo_n$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short 2>&1 | tail -40
for ch in 4; do
JWB_SLAB_CHUNKS=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n${N}_ch$ch.json 2> gpurun_out/r02_bench_c5_n${N}_ch$ch.err
tail -2 gpurun_out/r02_bench_c5_n${N}_ch$ch.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c5_n${N}_ch$ch.json").read().strip().splitlines()[-1])
    print("chunks $ch: value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), json.dumps(d.get("slab")), json.dumps(d.get("slab_parity")))
except Exception as e:
    print("FAILED", e)
PY
done
for ns in 1 2 8; do
JWB_SLAB_COPY_STREAMS=$ns timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5_n${N}_ns$ns.json 2> gpurun_out/r02_bench_c5_n${N}_ns$ns.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c5_n${N}_ns$ns.json").read().strip().splitlines()[-1])
    print("copy streams $ns: value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), json.dumps(d.get("slab")), d.get("slab_parity_max_err"))
except Exception as e:
    print("FAILED", e)
PY
done
JWB_SLAB_CHUNKS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 --slab-layout i > gpurun_out/r02_bench_c5_n${N}_layout_i.json 2> gpurun_out/r02_bench_c5_n${N}_layout_i.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c5_n${N}_layout_i.json").read().strip().splitlines()[-1])
    print("layout i: value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), json.dumps(d.get("slab")), d.get("slab_parity_max_err"))
except Exception as e:
    print("FAILED", e)
PY
