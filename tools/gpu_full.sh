#!/bin/bash
# full GPU suite + bench (no CPU baseline leg) ; logs in gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.log
