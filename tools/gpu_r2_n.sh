#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_conv.py -q -x -k "rowstats or folded" > gpurun_out/t_ln.log 2>&1; echo "ln tests rc=$?"; tail -3 gpurun_out/t_ln.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/t_gpu.log
for h in 1 0 1 0; do I360_LN_FOLD=$h timeout 900 python bench.py --no-cpu-baseline --no-side-configs --no-comparator > gpurun_out/bench_ln$h.log 2>gpurun_out/bench_ln$h.err; echo "bench fold=$h rc=$?"; tail -3 gpurun_out/bench_ln$h.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ln$h.log').read().strip().splitlines()[-1])
print('fold=$h ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'launches', d['gpu_launches'])
PY
cp gpurun_out/bench_breakdown.json gpurun_out/bench_breakdown_ln$h.json
done
