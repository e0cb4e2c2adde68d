#!/bin/bash
mkdir -p gpurun_out
python tools/vae_breakdown.py > gpurun_out/vae_breakdown.txt 2>&1; tail -45 gpurun_out/vae_breakdown.txt
