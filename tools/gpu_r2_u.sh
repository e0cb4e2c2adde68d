#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoders_gpu.py -q -x > gpurun_out/t_enc.log 2>&1; echo "encoder tests rc=$?"; tail -3 gpurun_out/t_enc.log
python tools/encoder_breakdown.py > gpurun_out/sam_breakdown.txt 2>&1; cat gpurun_out/sam_breakdown.txt | tail -8
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py 2>&1 | grep -E "ERROR SUMMARY|all new" 
