#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"; tail -3 gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/bench_2gpu.log
