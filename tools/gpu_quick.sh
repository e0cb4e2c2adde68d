#!/bin/bash
# Quick GPU visit: selected tests (-k "$1") and microbench sections ("$2"); logs in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$1" > gpurun_out/t_quick.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/t_quick.log
timeout 400 python tools/microbench.py $2 > gpurun_out/mb_quick.log 2>&1; echo "microbench rc=$?"; grep name gpurun_out/mb_quick.log
