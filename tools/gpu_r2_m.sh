#!/bin/bash
mkdir -p gpurun_out
for h in 1 0; do I360_CONV_HALO=$h timeout 300 python tools/microbench.py conv 2>&1 | grep -v cudnn | sed "s/^/halo=$h /" | cut -c1-150; done
