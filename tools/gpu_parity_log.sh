#!/bin/bash
# calibrated parity log (native vs fp32 oracle vs bf16 torch), printed by the tests themselves
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_calibrated_gpu.py tests/test_encoders_gpu.py -q -s > gpurun_out/parity_log.txt 2>&1; echo "rc=$?"; grep -E "native|passed|failed" gpurun_out/parity_log.txt | tail -40
