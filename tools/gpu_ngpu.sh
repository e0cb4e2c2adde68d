#!/bin/bash
# usage: gpu_ngpu.sh N  -> the bench line at N GPUs of one node (torchrun, one rank per GPU)
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "${N}gpu rc=$?"; tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e'], d.get('clocks'))
PY
