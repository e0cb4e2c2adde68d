"""Kernel micro-benchmarks (CUDA events, L2-flushed) -> gpurun_out/microbench.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagine360_b200 import ops

def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device="cuda")
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

res = []
def rec(name, ms, flops, bytes_):
    r = dict(name=name, ms=ms, tflops=flops / ms / 1e9, gbs=bytes_ / ms / 1e6)
    print(r, flush=True); res.append(r)

if "gemm" in sys.argv or len(sys.argv) == 1:
    for (M, N, K) in [(655360, 320, 320), (655360, 960, 320), (163840, 640, 640), (40960, 1280, 1280),
                      (262144, 320, 320), (655360, 2560, 320), (163840, 5120, 640), (40960, 10240, 1280),
                      (655360, 320, 1280), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
        act = ops.ACT_GEGLU if N in (2560, 5120, 10240) else 0
        out = torch.empty(M, N // 2 if act else N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm(a, w, act=act, out=out))
        rec(f"gemm {M}x{N}x{K} act{act}", ms, 2.0 * M * N * K, 2.0 * (M * K + N * K + out.numel()))
        ms = timeit(lambda: torch.matmul(a, w.t()))
        rec(f"  torch.matmul {M}x{N}x{K}", ms, 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N))
        del a, w, out
if "conv" in sys.argv or len(sys.argv) == 1:
    for (B, H, W, Ci, Co) in [(640, 32, 32, 320, 320), (640, 16, 16, 640, 640), (640, 8, 8, 1280, 1280),
                              (640, 4, 4, 1280, 1280), (32, 64, 132, 320, 320), (32, 32, 68, 640, 640),
                              (640, 32, 32, 960, 320), (1, 512, 1088, 128, 128), (1, 256, 544, 256, 256)]:
        x = torch.randn(B, H, W, Ci, device="cuda").bfloat16()
        w = torch.randn(Co, Ci, 3, 3, device="cuda").bfloat16()
        wp = ops.pack_conv3x3(w)
        ms = timeit(lambda: ops.conv3x3(x, wp))
        fl = 2.0 * B * H * W * 9 * Ci * Co
        rec(f"conv {B}x{H}x{W} {Ci}->{Co}", ms, fl, 2.0 * (x.numel() + wp.numel() + B * H * W * Co))
        xn = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wn = w.contiguous(memory_format=torch.channels_last)
        ms = timeit(lambda: torch.nn.functional.conv2d(xn, wn, padding=1))
        rec(f"  cudnn conv {B}x{H}x{W} {Ci}->{Co}", ms, fl, 0)
        del x, w, wp, xn, wn
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
