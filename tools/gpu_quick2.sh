#!/bin/bash
# usage: gpu_quick2.sh "<pytest args>" "<microbench sections>"
mkdir -p gpurun_out
timeout 600 python -m pytest $1 -x -q -m gpu > gpurun_out/t_quick.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/t_quick.log
if [ -n "$2" ]; then timeout 400 python tools/microbench.py $2 > gpurun_out/mb_quick.log 2>&1; echo "microbench rc=$?"; grep name gpurun_out/mb_quick.log; fi
