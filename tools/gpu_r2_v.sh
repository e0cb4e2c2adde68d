#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/t_gpu.log
for h in 1 0 1 0; do I360_PDL=$h timeout 900 python bench.py --no-cpu-baseline --no-comparator > gpurun_out/bench_pdl$h.log 2>gpurun_out/bench_pdl$h.err; echo "bench pdl=$h rc=$?"; tail -3 gpurun_out/bench_pdl$h.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl$h.log').read().strip().splitlines()[-1])
print('pdl=$h ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'c2', d['c2'].get('ms_per_step'), d['c2'].get('ms_per_step_eager'), 'c5', d['c5'].get('ms_per_step'), 'vae', d['vae_decode'].get('ms_per_frame'), 'c4', d['c4'].get('ms_per_clip'))
PY
done
