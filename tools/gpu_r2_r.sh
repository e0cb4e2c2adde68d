#!/bin/bash
bash tools/gpu_ncu.sh geglu_ln_k320 gemm_conv_kernel gemmln 655360 2560 320 1
bash tools/gpu_ncu.sh qkv_ln_k320 gemm_conv_kernel gemmln 655360 960 320 0
ls -la gpurun_out/*.ncu-rep
