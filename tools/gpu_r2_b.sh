#!/bin/bash
# round 2, run B: CUDA graph + 2-heads-per-CTA WarpAttn + attention cleanups
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -15 gpurun_out/t_gpu.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 5000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_breakdown.json'))
b=d['breakdown']
for k,v in b.items():
    if k!='shapes': print(k,v)
for k,v in b['shapes'].items():
    if k.startswith('attn'): print(v,k)
PY
