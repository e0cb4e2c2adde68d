"""Small instances of the kernel paths added after tools/sanitize_new_kernels.py was written, for compute-sanitizer
(memcheck / racecheck): sub-pixel upsample conv (+ GroupNorm statistics), stride-2 implicit conv (both pads, pano crop),
conv + per-group statistics (VAE), conv + per-channel statistics and the GroupNorm that folds them (UNet, opt-in), temporal
attention with two 16-frame tiles (F = 24 / 32 / 17).  Each result is also checked against torch so that a sanitizer-clean run
is a correct run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from imagine360_b200 import ops

BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)


def close(a, b, what, tol=2e-2):
    e = ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
    print(f"{what}: rel err {e:.4f}", flush=True)
    assert e < tol, what


def nchw(t):
    return t.float().permute(0, 3, 1, 2)



# nearest x2 upsample + conv3x3 in sub-pixel form, with and without the statistics epilogue
for (B, H, W, Ci, Co, crop) in [(1, 16, 16, 64, 128, 0), (2, 8, 12, 128, 160, 1)]:
    x, w, b = rn(B, H, W, Ci).to(BF), (rn(Co, Ci, 3, 3) / (9 * Ci) ** 0.5).to(BF), rn(Co).to(BF)
    weff = ops.pack_upsample_conv(w)
    up = F.interpolate(nchw(x), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), b.float(), padding=1)
    if crop:
        ref = ref[..., 2 * crop:-2 * crop]
    close(ops.conv_upsample2x(x, weff, b, crop=crop), ref.permute(0, 2, 3, 1), f"sub-pixel upsample conv crop={crop}")
    if crop == 0:
        out, st = ops.conv_upsample2x(x, weff, b, gn_groups=32)
        v = out.double().view(B, -1, 32, Co // 32)
        assert torch.allclose(st[..., 0], v.sum(dim=(1, 3)), rtol=1e-3, atol=0.5), "upsample conv group sums"
        print("sub-pixel upsample conv + group statistics OK", flush=True)

# stride-2 implicit conv: symmetric pad (UNet), asymmetric pad (VAE encoder), pano crop
for (B, H, W, Ci, Co, pad_lo, crop) in [(2, 16, 16, 64, 64, 1, 0), (1, 16, 24, 128, 128, 0, 0), (2, 16, 24, 64, 96, 1, 1)]:
    x, w, b = rn(B, H, W, Ci).to(BF), (rn(Co, Ci, 3, 3) / (9 * Ci) ** 0.5).to(BF), rn(Co).to(BF)
    if pad_lo:
        ref = F.conv2d(nchw(x), w.float(), b.float(), stride=2, padding=1)
    else:
        ref = F.conv2d(F.pad(nchw(x), (0, 1, 0, 1)), w.float(), b.float(), stride=2)
    if crop:
        ref = ref[..., crop:-crop]
    close(ops.conv3x3_s2(x, ops.pack_conv3x3(w), b, pad_lo=pad_lo, crop=crop), ref.permute(0, 2, 3, 1), f"stride-2 conv pad_lo={pad_lo} crop={crop}")

# conv + per-group statistics (group sizes 4 / 8 / 16) and conv + per-channel statistics (group sizes 10 / 20 / 40) -> GroupNorm
for (B, H, W, Ci, Co, kind) in [(2, 16, 24, 64, 128, "group"), (1, 64, 64, 128, 256, "group"), (3, 8, 8, 64, 320, "chan"), (5, 4, 8, 64, 640, "chan"),
                                (2, 16, 20, 64, 320, "chan_crop")]:
    crop = 2 if kind.endswith("crop") else 0
    x, w, b = rn(B, H, W, Ci).to(BF), (rn(Co, Ci, 3, 3) / (9 * Ci) ** 0.5).to(BF), rn(Co).to(BF)
    wp = ops.pack_conv3x3(w)
    gam, bet = (1 + 0.1 * rn(Co)).to(BF), (0.1 * rn(Co)).to(BF)
    if kind == "group":
        out, st = ops.conv3x3(x, wp, bias=b, gn_groups=32)
        y = ops.groupnorm(out, gam, bet, 32, 1e-6, True, stats=st)
    else:
        rv = rn(1, Co)
        out, st = ops.conv3x3(x, wp, bias=b, rowvec=rv, rowvec_div=B, crop=crop, chan_stats=True)
        v = out.double().view(B, -1, Co)
        assert torch.allclose(st[..., 0], v.sum(1), rtol=1e-4, atol=1e-2) and torch.allclose(st[..., 1], (v * v).sum(1), rtol=1e-4, atol=1e-2)
        y = ops.groupnorm(out, gam, bet, 32, 1e-5, True, chan_stats=st)
    ref = F.silu(F.group_norm(nchw(out), 32, gam.float(), bet.float(), 1e-6 if kind == "group" else 1e-5)).permute(0, 2, 3, 1)
    close(y, ref, f"conv + {kind} statistics -> GroupNorm Cout={Co}")

# temporal attention, two 16-frame tiles
for (B, Fr, D, heads, hd) in [(1, 24, 40, 8, 40), (2, 32, 9, 8, 80), (1, 17, 12, 8, 160)]:
    C = heads * hd
    qkv = rn(B * Fr * D, 3 * C).to(BF)
    out = torch.zeros(B * Fr * D, C, device="cuda", dtype=BF)
    ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, B, Fr, D, heads, hd)
    bd = lambda t: t.reshape(B, Fr, D, heads, hd).permute(0, 2, 3, 1, 4).float()
    ref = F.scaled_dot_product_attention(bd(qkv[:, :C]), bd(qkv[:, C:2 * C]), bd(qkv[:, 2 * C:]))
    close(bd(out), ref, f"temporal attention F={Fr} hd={hd}")
torch.cuda.synchronize()
print("all late-kernel instances OK")
