#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/t_gpu.log
for i in 1 2; do timeout 900 python bench.py --no-cpu-baseline --no-comparator --no-side-configs > gpurun_out/bench_rule$i.log 2>gpurun_out/bench_rule$i.err; tail -2 gpurun_out/bench_rule$i.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_rule$i.log').read().strip().splitlines()[-1])
print('run $i ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'launches', d['gpu_launches'])
b=json.load(open('gpurun_out/bench_breakdown.json'))['breakdown']
print({k:v for k,v in b.items() if k!='shapes' and v['ms']>3})
PY
done
