#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_conv.py -q -x -k "upsample or conv3x3" > gpurun_out/t_up.log 2>&1; echo "upsample tests rc=$?"; tail -12 gpurun_out/t_up.log
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_parity_calibrated_gpu.py tests/test_pipeline_gpu.py tests/test_full_size_gpu.py -q -x > gpurun_out/t_blk.log 2>&1; echo "block tests rc=$?"; tail -4 gpurun_out/t_blk.log
for i in 1 2; do for v in 1 0; do
  env I360_UPSAMPLE_SUBPIXEL=$v timeout 900 python bench.py --no-cpu-baseline --no-comparator > gpurun_out/bench_up$v.log 2>gpurun_out/bench_up$v.err; tail -2 gpurun_out/bench_up$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_up$v.log').read().strip().splitlines()[-1])
print('subpixel=$v ms_per_step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'c5', round(d['c5']['ms_per_step'],1), 'c2', round(d['c2']['ms_per_step'],2), 'vae ms/frame', round(d['vae_decode']['ms_per_frame'],3), 'c4', round(d['c4']['ms_per_clip'],1))
PY
done; done
