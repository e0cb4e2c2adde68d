#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --no-cpu-baseline --no-side-configs > gpurun_out/bench_cmp.log 2> gpurun_out/bench_cmp.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cmp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cmp.log').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'roofline', {k:d['roofline'][k] for k in ('achieved','frac','frac_of_per_launch_bound','per_launch_bound_note') if k in d['roofline']})
c=d['gpu_comparator']; print('comparator', c.get('ms_per_step'), c.get('ratio_vs_torch_cuda'), c.get('parity_vs_native'))
PY
