#!/bin/bash
# 2-GPU validation: multi-GPU tests, the default bench line at N = 2 (torchrun), the reference arm at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_env.txt
python -m pytest tests/test_gpu_multi.py tests/test_abi.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_n2_pytest_multi.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err ) 2> gpurun_out/r02_bench_n2.time
tail -3 gpurun_out/r02_n2_pytest_multi.txt gpurun_out/r02_bench_n2.time
tail -c 1500 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("c2", d["value"], d["n_gpus"], d.get("e2e"))
for k, v in d.get("workloads", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "error", "slab_parity_max_err", "scaling")}, (v.get("e2e") or {}).get("value"))
PY
