#!/bin/bash
# round 2, run A: new tests first (all failures shown), then the old suite, smoke, bench with every side block
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_calibrated_gpu.py tests/test_pipeline_call_gpu.py -q -m gpu -s > gpurun_out/t_new.log 2>&1; echo "new tests rc=$?"; grep -E "calibrated\]|passed|failed|Error|error" gpurun_out/t_new.log | tail -80
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_parity_calibrated_gpu.py --deselect tests/test_pipeline_call_gpu.py > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
