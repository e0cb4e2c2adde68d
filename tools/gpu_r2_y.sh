#!/bin/bash
mkdir -p gpurun_out
# ncu launch list of exactly ONE eager C3 step (cudaProfilerStart/Stop around it)
I360_PROFILE=1 I360_CUDA_GRAPH=0 timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-comparator --no-side-configs > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv; tail -2 gpurun_out/ncu_list.log | cut -c1-300
