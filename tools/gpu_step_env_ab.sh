#!/bin/bash
# generic env A/B of the C3 step: gpu_r2_ab.sh VAR A B  -> alternates VAR=A / VAR=B twice
mkdir -p gpurun_out
for i in 1 2; do for v in $2 $3; do
  env $1=$v timeout 900 python bench.py --no-cpu-baseline --no-comparator --no-side-configs > gpurun_out/bench_ab.log 2>gpurun_out/bench_ab.err; tail -2 gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ab.log').read().strip().splitlines()[-1])
b=json.load(open('gpurun_out/bench_breakdown.json'))['breakdown']
print('$1=$v ms_per_step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:v['ms'] for k,v in b.items() if k!='shapes' and ('norm' in k or 'temporal' in k)})
PY
done; done
