#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoders_gpu.py -q -x -s > gpurun_out/t_enc.log 2>&1; echo "encoder tests rc=$?"; grep -E "native|passed|failed|Error|error" gpurun_out/t_enc.log | tail -30
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/t_gpu.log
timeout 1200 python bench.py --no-cpu-baseline --no-side-configs --no-comparator > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; head -c 1500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
