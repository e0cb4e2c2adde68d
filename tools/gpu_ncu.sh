#!/bin/bash
# ncu --set full of one kernel shape: usage gpu_ncu.sh <outname> <kernel-regex> <one_kernel.py args...>
mkdir -p gpurun_out
out=$1; rx=$2; shift 2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/$out -f python tools/one_kernel.py "$@" > gpurun_out/ncu_$out.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$out.log
