#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_conv.py -q -m gpu -k "folded" > gpurun_out/t_ln.log 2>&1; echo "gemm_ln tests rc=$?"; tail -15 gpurun_out/t_ln.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -8 gpurun_out/t_gpu.log
timeout 1200 python bench.py --no-cpu-baseline --no-side-configs --no-comparator > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
I360_LN_FOLD=0 timeout 1200 python bench.py --no-cpu-baseline --no-side-configs --no-comparator > gpurun_out/bench_nofold.log 2>gpurun_out/bench_nofold.err; echo "bench(no fold) rc=$?"; head -c 400 gpurun_out/bench_nofold.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_breakdown.json'))
b=d['breakdown']
for k,v in b.items():
    if k!='shapes': print(k,v)
PY
