"""Per-launch device time of one SAM ViT-B batch (8 frames at 1024 px) through the C ABI, grouped by entry point and shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagine360_b200 import ops
from imagine360_b200.host.encoders import SamImageEncoderNative
from imagine360_b200.host.sam import build_sam_vit_b

torch.set_grad_enabled(False)
sam = build_sam_vit_b()
enc = sam.image_encoder.cuda()
nat = SamImageEncoderNative(enc)
x = torch.randn(8, 3, 1024, 1024, device="cuda")
for _ in range(2):
    nat(x)
torch.cuda.synchronize()
L = ops.lib()
recs = []
class T:
    def __init__(s, n, f): s.n, s.f = n, f
    def __call__(s, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.f(*a); e1.record(); recs.append((s.n, e0, e1)); return r
class P:
    def __getattr__(s, n):
        f = getattr(L, n); return T(n, f) if n.startswith("i360_") and callable(f) and "slots" not in n and "supported" not in n else f
px = P(); ol = ops.lib
ops.lib = lambda: px
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record(); nat(x); s1.record(); torch.cuda.synchronize()
ops.lib = ol
agg = {}
for n, a, b in recs:
    c, ms = agg.get(n, (0, 0.0)); agg[n] = (c + 1, ms + a.elapsed_time(b))
tot = sum(v[1] for v in agg.values())
print(f"SAM ViT-B, 8 frames: {s0.elapsed_time(s1):.2f} ms wall, {tot:.2f} ms inside C-ABI launches, {len(recs)} launches")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{c:3d}  {n}")
