#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_conv.py -q -x > gpurun_out/t_gc.log 2>&1; echo "gemm/conv tests rc=$?"; tail -6 gpurun_out/t_gc.log
timeout 300 python tools/ln_fold_probe.py 2>&1 | grep -v "^{" | tee gpurun_out/ln_fold_probe_ng4.txt | tail -8
for h in 4 2 4 2; do I360_EPI_GROUPS=$h timeout 900 python bench.py --no-cpu-baseline --no-side-configs --no-comparator > gpurun_out/bench_ng$h.log 2>gpurun_out/bench_ng$h.err; echo "bench groups=$h rc=$?"; tail -3 gpurun_out/bench_ng$h.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ng$h.log').read().strip().splitlines()[-1])
print('groups=$h ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])
PY
cp gpurun_out/bench_breakdown.json gpurun_out/bench_breakdown_ng$h.json
done
