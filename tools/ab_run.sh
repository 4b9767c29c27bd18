# scratch driver for same-run A/B sweeps on the GPU box (edited per experiment)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_tail.log
for v in base ""; do
  lib=/root/repo/jwave_b200/libjwave_cuda${v:+_$v}.so
  echo "== $lib" >> gpurun_out/ab_tail.log
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c3 "" "" wpt_threads=288 wpt_r=4,wpt_rs=4 >> gpurun_out/ab_tail.log 2>&1
done
python tools/sweep.py c2 rev_tile=4096 rev_tile=4096,rev_m=5 rev_tile=4096,rev_m=3 rev_tile=8192 rev_tile=4096,rev_threads=160 rev_tile=4096,rev_threads=96 rev_tile=4096,fwd_tile=4096 >> gpurun_out/ab_tail.log 2>&1
cat gpurun_out/pytest_tail.log gpurun_out/ab_tail.log
