"""Bandwidth of the memory-bound kernels vs a plain copy, standalone and behind a tensor-heavy kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagine360_b200 import ops

def t(fn, iters=20, pre=None):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        if pre is not None: pre()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

flush = torch.empty(256 << 20, dtype=torch.int8, device="cuda")
A = torch.randn(8192, 8192, device="cuda").bfloat16()
def heat():
    for _ in range(6): torch.matmul(A, A)
for (M, C) in [(655360, 320), (163840, 640), (262144, 320), (40960, 1280)]:
    x = torch.randn(M, C, device="cuda").bfloat16(); y = torch.empty_like(x)
    g = torch.ones(C, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda").bfloat16()
    by = 4.0 * M * C
    for nm, pre in (("flush", lambda: flush.zero_()), ("heat", heat), ("none", None)):
        c = t(lambda: y.copy_(x), pre=pre)
        l = t(lambda: ops.layernorm(x, g, b, out=y) if "out" in ops.layernorm.__code__.co_varnames else ops.layernorm(x, g, b), pre=pre)
        print(f"M={M} C={C} pre={nm}: copy {c:.3f} ms {by/c/1e9:.2f} TB/s | LN {l:.3f} ms {by/l/1e9:.2f} TB/s", flush=True)
    B = M // 1024
    x4 = x.view(B, 32, 32, C)
    for nm, pre in (("flush", lambda: flush.zero_()), ("heat", heat)):
        gn = t(lambda: ops.groupnorm(x4, g, b, 32, 1e-5, True), pre=pre)
        print(f"   GN stats+apply pre={nm}: {gn:.3f} ms {1.5*by/gn/1e9:.2f} TB/s (3 passes)", flush=True)
    del x, y, x4
