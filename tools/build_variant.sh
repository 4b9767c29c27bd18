#!/bin/bash
# Builds an A/B variant of libjwave_cuda.so: the listed sources recompiled with extra flags, the other objects taken
# from the regular build.  The variant lands in variants/<name>/libjwave_cuda.so (git-ignored, travels with gpurun);
# run it with  LD_LIBRARY_PATH=variants/<name> tools/qbench ...
#   tools/build_variant.sh <name> "<extra nvcc flags>" file.cu [file.cu ...]
set -e
name=$1; flags=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root/jwave_b200/csrc"
mkdir -p "$root/variants/$name/obj"
objs=""
for o in build/*.o; do
  b=$(basename $o .o)
  if [[ " $* " == *" $b.cu "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include $flags -c $b.cu -o "$root/variants/$name/obj/$b.o" &
    objs="$objs $root/variants/$name/obj/$b.o"
  else
    objs="$objs $o"
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/variants/$name/libjwave_cuda.so" $objs -cudart static
ls -la "$root/variants/$name/libjwave_cuda.so"
