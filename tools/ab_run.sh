# scratch driver for GPU-box experiments (edited per experiment)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wpt or WPT or packet" 2>&1 | tail -3 > gpurun_out/pytest_tail.log
for v in base ""; do
  lib=/root/repo/jwave_b200/libjwave_cuda${v:+_$v}.so
  echo "== $lib" >> gpurun_out/ab_rot.log
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c3 "" "" >> gpurun_out/ab_rot.log 2>&1
done
cat gpurun_out/pytest_tail.log gpurun_out/ab_rot.log
