#!/bin/bash
# usage: tools/nb_variants.sh build "<defs 1>" "<defs 2>" ...   (anywhere: nvcc cross-compiles) builds each variant of the
#        library into build/variants/ (git-ignored, travels to the GPU box);
#        tools/nb_variants.sh run                                  (on the GPU box) prints tools/nb_time.py's line per variant.
# NB_TIME_CASES (default murd64) selects the systems.
d=build/variants
if [ "$1" = build ]; then
  shift; rm -rf $d; mkdir -p $d; k=0
  for defs in "$@"; do
    echo "$defs" > $d/defs_$k.txt
    ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $defs -o $d/lib_$k.so \
         molearn_b200/csrc/energy.cu molearn_b200/csrc/topology.cpp > $d/build_$k.log 2>&1 || echo "build failed: $defs" ) &
    k=$((k+1))
  done
  wait
else
  export NB_TIME_CASES=${NB_TIME_CASES:-murd64}
  for f in $(ls $d/defs_*.txt | sort -t_ -k2 -n); do
    k=${f##*_}; k=${k%.txt}
    [ -f $d/lib_$k.so ] && MOLEARN_B200_LIB=$d/lib_$k.so python ${NB_TIME_SCRIPT:-tools/nb_time.py} "$(cat $f)" 2>&1 | tail -1
  done
fi
