#!/bin/bash
mkdir -p gpurun_out
python tools/sanitize_new_kernels.py > gpurun_out/new_kernels_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/new_kernels_plain.log
for tool in memcheck racecheck; do
  for grp in 2 4; do
    I360_EPI_GROUPS=$grp timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_new_kernels.py > gpurun_out/sanitizer_${tool}_groups$grp.log 2>&1
    echo "$tool groups=$grp rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all new-kernel" gpurun_out/sanitizer_${tool}_groups$grp.log | tail -3
  done
done
