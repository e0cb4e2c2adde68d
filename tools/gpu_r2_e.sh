#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_full_size_gpu.py tests/test_parity_calibrated_gpu.py -q -m gpu -k "attention or attn" > gpurun_out/t_attn.log 2>&1; echo "attn tests rc=$?"; tail -6 gpurun_out/t_attn.log
timeout 300 python tools/attn_ab.py > gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
bash tools/gpu_ncu.sh attn2_pano attention2_kernel attn 8 8192 5 64
