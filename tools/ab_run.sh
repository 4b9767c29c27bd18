# scratch driver for GPU-box experiments (edited per experiment)
for v in base ""; do
  lib=/root/repo/jwave_b200/libjwave_cuda${v:+_$v}.so
  echo "== $lib" >> gpurun_out/ab_strcap.log
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c4 "" str_rev_tile=256 str_tile=256,str_rev_tile=256 str_threads=256,str_rev_threads=256 >> gpurun_out/ab_strcap.log 2>&1
  JWAVE_CUDA_LIB=$lib python tools/sweep.py c5 "" >> gpurun_out/ab_strcap.log 2>&1
done
cat gpurun_out/ab_strcap.log
