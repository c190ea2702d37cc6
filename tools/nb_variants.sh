#!/bin/bash
# usage: tools/nb_variants.sh "<defs 1>" "<defs 2>" ...   builds each variant of the library next to the shipped one and
# prints tools/nb_time.py's line for it (run on the GPU box)
mkdir -p gpurun_out/variants
k=0
for defs in "$@"; do
  lib=gpurun_out/variants/lib_$k.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $defs -o $lib \
       molearn_b200/csrc/energy.cu molearn_b200/csrc/topology.cpp > gpurun_out/variants/build_$k.log 2>&1 || { echo "build failed: $defs"; tail -5 gpurun_out/variants/build_$k.log; k=$((k+1)); continue; }
  MOLEARN_B200_LIB=$lib python tools/nb_time.py "$defs" 2>&1 | tail -1
  k=$((k+1))
done
