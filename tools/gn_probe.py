"""GroupNorm stats / apply timed separately (CUDA events) + one plain call for ncu."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagine360_b200 import ops
from imagine360_b200.ops import lib, _p, _stream, c_int, c_float

def t(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.int8, device="cuda")
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

shapes = [(640, 32, 32, 320, 0), (32, 64, 128, 320, 2), (640, 16, 16, 640, 0), (640, 32, 32, 640, 0)]
if "once" in sys.argv: shapes = shapes[:2]
for (B, H, W, C, pad) in shapes:
    x = torch.randn(B, H, W, C, device="cuda").bfloat16()
    g = torch.ones(C, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda").bfloat16()
    stats = torch.empty(B, 32, 2, dtype=torch.float64, device="cuda")
    out = torch.empty(B, H, W + 2 * pad, C, device="cuda", dtype=torch.bfloat16)
    st = lambda: lib().i360_groupnorm_stats(_p(x), c_int(C), None, c_int(0), c_int(B), c_int(H), c_int(W), c_int(pad), c_int(32), _p(stats), _stream())
    ap = lambda silu: lib().i360_groupnorm_apply(_p(x), c_int(C), None, c_int(0), c_int(B), c_int(H), c_int(W), c_int(pad), c_int(32), _p(stats),
                                                 ctypes.c_double(0.0), _p(g), _p(b), c_float(1e-5), c_int(silu), _p(out), _stream())
    if "once" in sys.argv:
        st(); ap(1); torch.cuda.synchronize(); continue
    by = 2.0 * x.numel()
    a, b1, b0 = t(st), t(lambda: ap(1)), t(lambda: ap(0))
    print(f"{B}x{H}x{W}x{C} pad{pad}: stats {a:.3f} ms {by/a/1e9:.2f} TB/s | apply+silu {b1:.3f} ms {(by+2.0*out.numel())/b1/1e9:.2f} TB/s | apply {b0:.3f} ms", flush=True)
