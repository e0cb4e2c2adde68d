#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -6 gpurun_out/t_gpu.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.log
