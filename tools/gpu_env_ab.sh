#!/bin/bash
# A/B by environment variable on one box: usage gpu_env_ab.sh "<VAR=value>" "<microbench sections>"
mkdir -p gpurun_out
for rep in 1 2; do
  for which in base exp; do
    if [ $which = exp ]; then export $1; else unset ${1%%=*}; fi
    timeout 300 python tools/microbench.py $2 > gpurun_out/ab_${which}_$rep.log 2>&1
    echo "== $which $rep rc=$?"; grep name gpurun_out/ab_${which}_$rep.log | grep -v "torch\|plain" | sed 's/, .tflops.: /  TF /; s/, .gbs.*//'
  done
done
