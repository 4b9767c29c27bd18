# scratch driver for GPU-box runs (edited per experiment)
for v in base sel contig; do
  lib=/root/repo/jwave_b200/libjwave_cuda_$v.so
  echo "== $lib" >> gpurun_out/ab_taps2.log
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c4 "" >> gpurun_out/ab_taps2.log 2>&1
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c5 "" >> gpurun_out/ab_taps2.log 2>&1
done
JWAVE_CUDA_LIB=/root/repo/jwave_b200/libjwave_cuda_contig.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/ab_taps2.log
cat gpurun_out/ab_taps2.log
