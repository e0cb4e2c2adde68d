#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/attn_ab.log
timeout 200 python tools/attn_ab.py quick >> gpurun_out/attn_ab.log 2>&1
for v in nodelay nodelay_spin delay_spin; do
  I360_LIB_PATH=$PWD/tools/_ab/lib_$v.so timeout 200 python tools/attn_ab.py quick >> gpurun_out/attn_ab.log 2>&1
done
I360_ATTN_V2=0 timeout 200 python tools/attn_ab.py quick >> gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
