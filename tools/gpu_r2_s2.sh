#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_conv.py -q -x -k "stride2" > gpurun_out/t_s2.log 2>&1; echo "stride2 tests rc=$?"; tail -12 gpurun_out/t_s2.log
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_parity_calibrated_gpu.py tests/test_pipeline_gpu.py tests/test_full_size_gpu.py tests/test_pipeline_call_gpu.py -q -x > gpurun_out/t_blk.log 2>&1; echo "block tests rc=$?"; tail -4 gpurun_out/t_blk.log
for i in 1 2; do for v in 0 1; do
  env I360_CONV_S2_IM2COL=$v timeout 900 python bench.py --no-cpu-baseline --no-comparator --no-side-configs > gpurun_out/bench_s2$v.log 2>gpurun_out/bench_s2$v.err; tail -2 gpurun_out/bench_s2$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s2$v.log').read().strip().splitlines()[-1])
print('im2col=$v ms_per_step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])
PY
done; done
