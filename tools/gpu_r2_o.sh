#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ln_fold_probe.py 2>&1 | grep -v "^{" | tee gpurun_out/ln_fold_probe.txt
