#!/bin/bash
# A/B on one box: tools/_ab/lib_old.so vs the in-tree library; usage: gpu_ab.sh "<microbench sections>"
mkdir -p gpurun_out
for rep in 1 2; do
  for which in old new; do
    if [ $which = old ]; then export I360_LIB_PATH=$PWD/tools/_ab/lib_old.so; else unset I360_LIB_PATH; fi
    timeout 300 python tools/microbench.py $1 > gpurun_out/ab_${which}_$rep.log 2>&1
    echo "== $which $rep rc=$?"; grep name gpurun_out/ab_${which}_$rep.log | grep -v torch | sed 's/, .tflops.*//'
  done
done
