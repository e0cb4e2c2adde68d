#!/bin/bash
# sweep one environment knob over values on one box: gpu_sweep.sh VAR "v1 v2 ..." "<microbench sections>" "<grep filter>"
mkdir -p gpurun_out
for v in $2; do
  export $1=$v
  timeout 200 python tools/microbench.py $3 > gpurun_out/sweep.log 2>&1
  echo "== $1=$v rc=$?"; grep name gpurun_out/sweep.log | grep "$4" | sed 's/, .tflops.*//'
done
