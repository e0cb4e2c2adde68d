#!/bin/bash
# round-end record: GPU suite, smoke, the default bench line (all blocks); logs in gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
t0=$(date +%s); timeout 1800 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline'])
print('comparator', d['gpu_comparator'].get('ms_per_step'), d['gpu_comparator'].get('ratio_vs_torch_cuda'))
print('c2', d['c2'].get('ms_per_step'), 'c5', d['c5'].get('ms_per_step'), 'c4', d['c4'].get('ms_per_clip'), 'vae', d['vae_decode'].get('ms_per_frame'), d['vae_decode'].get('tensor_frac'))
print('encoders', d['encoders'])
PY
