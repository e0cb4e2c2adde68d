#!/bin/bash
# full GPU suite + smoke + bench (with the CPU baseline leg) ; logs in gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py $1 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/bench.log
