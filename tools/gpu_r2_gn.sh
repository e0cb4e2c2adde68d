#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_conv.py -q -x -k "groupnorm_statistics" > gpurun_out/t_gn.log 2>&1; echo "gn-stats tests rc=$?"; tail -12 gpurun_out/t_gn.log
timeout 900 python -m pytest tests -q -x -m gpu -k "vae or pipeline_call or decode" > gpurun_out/t_vae.log 2>&1; echo "vae tests rc=$?"; tail -4 gpurun_out/t_vae.log
for v in 1 0 1 0; do I360_VAE_GN_FUSED=$v python tools/vae_breakdown.py 2>&1 | head -4 | sed "s/^/fused=$v /"; done
