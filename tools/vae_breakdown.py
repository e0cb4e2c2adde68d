"""Per-launch device time of one AutoencoderKL.decode call (4 frames of a padded 16x512x1024 latent: 64 x 136) through the
C ABI, CUDA events around every launch; prints the launches grouped by (entry point, shape)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagine360_b200 import ops
from imagine360_b200.host.vae import AutoencoderKL

torch.set_grad_enabled(False)
dev = "cuda"
vae = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                    block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, norm_num_groups=32).to(dev, torch.bfloat16)
frames = int(os.environ.get("FRAMES", 4))
z = torch.randn(frames, 4, 64, 136, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    vae.decode(z)
torch.cuda.synchronize()
L = ops.lib()
recs = []
og, oc = ops.gemm, ops.conv3x3
tag = [None]
def gemm(a, w, *args, **kw):
    tag[0] = f"gemm M={a.shape[0]} N={w.shape[0]} K={a.shape[1]}"; tag.append(2.0 * a.shape[0] * w.shape[0] * a.shape[1]); r = og(a, w, *args, **kw); return r
def conv(x, wp, *args, **kw):
    b, h, wd, _ = x.shape
    tag[0] = f"conv {b}x{h}x{wd} Cout={wp.shape[0]} K={wp.shape[1]}"; tag.append(2.0 * b * h * wd * wp.shape[0] * wp.shape[1]); return oc(x, wp, *args, **kw)
class T:
    def __init__(s, n, f): s.n, s.f = n, f
    def __call__(s, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.f(*a); e1.record()
        fl = tag.pop() if len(tag) > 1 else 0.0
        recs.append((s.n, tag[0] if s.n in ("i360_gemm_bf16", "i360_conv3x3_bf16") else "", fl, e0, e1)); return r
class P:
    def __getattr__(s, n):
        f = getattr(L, n); return T(n, f) if n.startswith("i360_") and callable(f) else f
px = P(); ol = ops.lib
ops.lib, ops.gemm, ops.conv3x3 = (lambda: px), gemm, conv
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record(); vae.decode(z); s1.record(); torch.cuda.synchronize()
ops.lib, ops.gemm, ops.conv3x3 = ol, og, oc
agg = {}
for n, t, fl, a, b in recs:
    k = (n, t); c, ms, f = agg.get(k, (0, 0.0, 0.0)); agg[k] = (c + 1, ms + a.elapsed_time(b), f + fl)
tot = sum(v[1] for v in agg.values())
print(f"decode {frames} frames: {s0.elapsed_time(s1):.2f} ms wall, {tot:.2f} ms in kernels, {len(recs)} launches")
for (n, t), (c, ms, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{c:3d}  {n} {t}  {f/ms/1e9 if f else 0:7.1f} TF/s")
