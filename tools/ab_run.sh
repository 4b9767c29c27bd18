# scratch driver for GPU-box runs (edited per experiment)
N=${NGPU:-2}
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_multi.log
cat gpurun_out/pytest_multi.log
for mode in ${MODES:-copies}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 --no-cpu --no-e2e --slab $mode > gpurun_out/bench_c5_n${N}_$mode.json 2> gpurun_out/bench_c5_n${N}_$mode.err
  tail -3 gpurun_out/bench_c5_n${N}_$mode.err | grep -v "OMP_NUM\|\*\*\*"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c5_n${N}_$mode.json').read().strip().splitlines()[-1])
    print('$mode', d['value'], d['forward_gsps'], d['reverse_gsps'], d['roundtrip_max_abs_err'], d['config']['workload'][-80:])
except Exception as e: print('$mode FAILED', e)
PY
done
