#!/bin/bash
mkdir -p gpurun_out
for p in 0 1; do PROBE_PASS=$p timeout 60 tools/_ab/probe_halo_desc > gpurun_out/probe_halo_$p.log 2>&1; echo "probe pass $p rc=$?"; cut -c1-200 gpurun_out/probe_halo_$p.log | head -10; done
