#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/_ab/probe_ts_mma > gpurun_out/probe_ts.log 2>&1; echo "probe rc=$?"; head -12 gpurun_out/probe_ts.log | cut -c1-400
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -8 gpurun_out/t_gpu.log
timeout 1200 python bench.py --no-cpu-baseline --no-side-configs > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_breakdown.json'))
b=d['breakdown']
for k,v in b.items():
    if k!='shapes': print(k,v)
for k,v in b['shapes'].items():
    if k.startswith('attn'): print(v,k)
PY
