#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1800 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/bench_full.err; tail -c 7000 gpurun_out/bench_full.log
