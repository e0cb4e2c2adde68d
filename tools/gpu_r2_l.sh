#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_conv.py -q -x -k "conv" > gpurun_out/t_conv.log 2>&1; echo "conv tests rc=$?"; tail -12 gpurun_out/t_conv.log
timeout 300 python tools/vae_breakdown.py > gpurun_out/vae_breakdown_halo.txt 2>&1; head -12 gpurun_out/vae_breakdown_halo.txt
I360_CONV_HALO=0 timeout 300 python tools/vae_breakdown.py > gpurun_out/vae_breakdown_nohalo.txt 2>&1; head -3 gpurun_out/vae_breakdown_nohalo.txt
for h in 1 0 1 0; do I360_CONV_HALO=$h timeout 900 python bench.py --no-cpu-baseline --no-side-configs --no-comparator > gpurun_out/bench_halo$h.log 2>gpurun_out/bench_halo$h.err; echo "bench halo=$h rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_halo$h.log').read().strip().splitlines()[-1])
print('halo=$h ms_per_step', d['ms_per_step'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])
b=json.load(open('gpurun_out/bench_breakdown.json'))['breakdown']
print({k:v for k,v in b.items() if k!='shapes' and v['ms']>3})
PY
done
