# scratch driver for GPU-box runs (edited per experiment)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c4 --batch 4 --no-cpu --no-e2e > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --workload c5 --no-cpu --no-e2e > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
tail -c 600 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; for f in default ref c3 c4 c5; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
print('$f', d.get('value'), d.get('unit'), d.get('ms_per_step'), d.get('roofline',{}).get('frac'), d.get('forward_gsps'), d.get('reverse_gsps'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'))
PY
done
