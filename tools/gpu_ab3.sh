#!/bin/bash
# A/B/C on one box: libs given as args (first = in-tree when "-"); usage: gpu_ab3.sh "<sections>" lib1 lib2 ...
mkdir -p gpurun_out
sec=$1; shift
for rep in 1 2; do
  for lib in "$@"; do
    if [ "$lib" = "-" ]; then unset I360_LIB_PATH; else export I360_LIB_PATH=$PWD/$lib; fi
    timeout 300 python tools/microbench.py $sec > gpurun_out/ab3.log 2>&1
    echo "== $lib rep $rep rc=$?"; grep name gpurun_out/ab3.log | grep -v torch | sed 's/, .tflops.*//'
  done
done
