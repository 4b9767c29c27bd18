set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_swz.log
for v in base swz ""; do
  lib=/root/repo/jwave_b200/libjwave_cuda${v:+_$v}.so
  echo "== $lib" >> gpurun_out/ab_swz.log
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c2 "" "" >> gpurun_out/ab_swz.log 2>&1
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c4 "" >> gpurun_out/ab_swz.log 2>&1
done
cat gpurun_out/pytest_swz.log gpurun_out/ab_swz.log
