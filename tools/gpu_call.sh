#!/bin/bash
# One GPU-box visit: new-kernel tests -> microbench -> ncu of the new kernel -> full GPU suite -> bench.  Logs in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "cross_attention" > gpurun_out/t_xattn.log 2>&1
echo "xattn tests rc=$?"; tail -5 gpurun_out/t_xattn.log
timeout 300 python tools/microbench.py xfused xattn > gpurun_out/mb_xattn.log 2>&1; echo "microbench rc=$?"; grep xfused gpurun_out/mb_xattn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xattn_kernel -c 1 -o gpurun_out/xattn_full -f python tools/microbench.py xfused > gpurun_out/ncu_xattn.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/t_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log
