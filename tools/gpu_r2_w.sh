#!/bin/bash
mkdir -p gpurun_out
run() { # name env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --no-cpu-baseline --no-comparator --no-side-configs > gpurun_out/bench_$name.log 2>gpurun_out/bench_$name.err; tail -2 gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$name.log').read().strip().splitlines()[-1])
print('$name ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
PY
}
for i in 1 2; do
run notrigger I360_LIB_PATH=$PWD/tools/_ab/lib_pdl_notrigger.so I360_PDL=1
run early I360_PDL=1
run off I360_PDL=0
done
