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

def timeit_hot(fn, heat, n=20):
    """average launch time inside a long busy stretch (sustained clocks): `heat` keeps the GPU busy ~0.5 s first"""
    for _ in range(heat): fn()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

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
if "gemmr" in sys.argv:
    # the in-step flavour of the projection GEMMs: bias + residual epilogue
    shapes = [(655360, 320, 320), (262144, 320, 320), (163840, 640, 640), (65536, 640, 640), (40960, 1280, 1280), (655360, 320, 1280), (163840, 640, 2560)]
    if "gemmq" in sys.argv:     # QKV / wide projections (tile-width experiments)
        shapes = [(655360, 960, 320), (262144, 960, 320), (163840, 1920, 640), (65536, 1920, 640), (163840, 640, 640), (65536, 640, 640), (163840, 640, 2560)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
        b = torch.randn(N, device="cuda").bfloat16(); r = torch.randn(M, N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for nm, fn in (("plain", lambda: ops.gemm(a, w, out=out)), ("bias", lambda: ops.gemm(a, w, bias=b, out=out)),
                       ("bias+resid", lambda: ops.gemm(a, w, bias=b, resid=r, out=out))):
            by = 2.0 * (M * K + N * K + M * N * (2 if "resid" in nm else 1))
            ms = timeit(fn)
            rec(f"gemm {M}x{N}x{K} {nm}", ms, 2.0 * M * N * K, by)
            rec(f"gemm {M}x{N}x{K} {nm} HOT", timeit_hot(fn, int(300 / ms)), 2.0 * M * N * K, by)
        del a, w, b, r, out
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
if "attn" in sys.argv or len(sys.argv) == 1:
    for (imgs, N, heads, hd) in [(32, 8192, 5, 64), (640, 1024, 5, 64), (32, 2048, 10, 64), (640, 256, 10, 64), (32, 512, 20, 64), (640, 64, 20, 64)]:
        C = heads * hd
        qkv = torch.randn(imgs * N, 3 * C, device="cuda").bfloat16()
        out = torch.empty(imgs * N, C, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.attention(ops.seq_view(qkv, imgs, N, 0), ops.seq_view(qkv, imgs, N, C), ops.seq_view(qkv, imgs, N, 2 * C),
                                   ops.seq_view(out, imgs, N), heads, hd, imgs)
        ms = timeit(fn)
        fl = 4.0 * imgs * heads * N * N * hd
        rec(f"attn self imgs{imgs} N{N} h{heads} d{hd}", ms, fl, 2.0 * 4 * imgs * N * C)
        if "sdpa" in sys.argv:
            q4 = qkv.view(imgs, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
            ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4[0], q4[1], q4[2]))
            rec(f"  torch sdpa imgs{imgs} N{N}", ms, fl, 0)
            del q4
        del qkv, out
    # WarpAttn level enc0: equi 2048 tokens <- 20 views x 256, heads 10 (C=320), hd 32, bias
    b, Fr, m, hw, EN, heads, hd = 2, 16, 20, 256, 2048, 10, 32
    C = heads * hd
    pers_kv = torch.randn(b * m * Fr * hw, 2 * C, device="cuda").bfloat16()
    equi = torch.randn(b * Fr * EN, C, device="cuda").bfloat16()
    bias = torch.randn(EN, m * hw, device="cuda").bfloat16()
    out = torch.empty_like(equi)
    fn = lambda: ops.attention(ops.seq_view(equi, b * Fr, EN), ops.multiview_view(pers_kv, b, m, Fr, hw, 0),
                               ops.multiview_view(pers_kv, b, m, Fr, hw, C), ops.seq_view(out, b * Fr, EN), heads, hd, b * Fr, bias=bias)
    rec("attn warp equi<-pers enc0", timeit(fn), 4.0 * b * Fr * heads * EN * m * hw * hd, 0)
if "xattn" in sys.argv or len(sys.argv) == 1:
    # text / image-prompt cross attention: K/V of a clip element shared by its 16 frames
    for (imgs, N, heads, hd, nk) in [(640, 1024, 5, 64, 77), (640, 1024, 5, 64, 64), (32, 8192, 5, 64, 77), (640, 256, 10, 64, 77)]:
        C = heads * hd; fr = 16
        q = torch.randn(imgs * N, C, device="cuda").bfloat16()
        kv = torch.randn((imgs // fr) * nk, 2 * C, device="cuda").bfloat16()
        out = torch.zeros(imgs * N, C, device="cuda", dtype=torch.bfloat16)
        for acc in (False, True):
            fn = lambda: ops.attention(ops.seq_view(q, imgs, N), ops.seq_view(kv, imgs // fr, nk, 0, share_div=fr),
                                       ops.seq_view(kv, imgs // fr, nk, C, share_div=fr), ops.seq_view(out, imgs, N), heads, hd, imgs, accumulate=acc)
            by = 2.0 * (q.numel() + out.numel() * (2 if acc else 1))
            rec(f"xattn imgs{imgs} N{N} h{heads} nk{nk} acc{int(acc)}", timeit(fn), 4.0 * imgs * heads * N * nk * hd, by)
        del q, kv, out
if "xfused" in sys.argv or len(sys.argv) == 1:
    # fused text + image-prompt cross attention with stationary K/V (every attn2 of the C3 step)
    for (imgs, N, heads) in [(640, 1024, 5), (32, 8192, 5), (640, 256, 10), (32, 2048, 10), (640, 64, 20), (32, 512, 20), (640, 16, 20), (32, 128, 20)]:
        C = heads * 64; fr = 16
        q = torch.randn(imgs * N, C, device="cuda").bfloat16()
        kvt = torch.randn((imgs // fr) * 77, 2 * C, device="cuda").bfloat16()
        kvi = torch.randn((imgs // fr) * 64, 2 * C, device="cuda").bfloat16()
        out = torch.zeros(imgs * N, C, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.cross_attention_text_ip(q, out, kvt, 77, kvi, 64, imgs // fr, heads, 64)
        rec(f"xfused imgs{imgs} N{N} h{heads}", timeit(fn), 4.0 * imgs * heads * N * 141 * 64, 2.0 * (q.numel() + out.numel()))
        del q, kvt, kvi, out
if "norm" in sys.argv or len(sys.argv) == 1:
    for (B, H, W, C) in [(640, 32, 32, 320), (32, 64, 128, 320), (640, 16, 16, 640), (640, 8, 8, 1280)]:
        x = torch.randn(B, H, W, C, device="cuda").bfloat16()
        g = torch.ones(C, device="cuda").bfloat16(); bta = torch.zeros(C, device="cuda").bfloat16()
        ms = timeit(lambda: ops.groupnorm(x, g, bta, 32, 1e-5, True))
        rec(f"groupnorm+silu {B}x{H}x{W}x{C}", ms, 0, 2.0 * 3 * x.numel())
        ms = timeit(lambda: ops.layernorm(x.view(-1, C), g, bta))
        rec(f"layernorm {B*H*W}x{C}", ms, 0, 2.0 * 2 * x.numel())
        del x
if "temporal" in sys.argv:
    # VersatileAttention over the frame axis: rows (b, f, d), q/k/v column slices of the fused projection
    for (B, D, heads, hd) in [(40, 1024, 8, 40), (2, 8192, 8, 40), (40, 256, 8, 80), (2, 2048, 8, 80), (40, 64, 8, 160), (2, 512, 8, 160)]:
        C = heads * hd; Fr = 16
        qkv = torch.randn(B * Fr * D, 3 * C, device="cuda").bfloat16()
        out = torch.empty(B * Fr * D, C, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, B, Fr, D, heads, hd)
        rec(f"temporal B{B} D{D} C{C}", timeit(fn), 4.0 * B * D * heads * Fr * Fr * hd, 2.0 * 4 * B * Fr * D * C)
        del qkv, out
if "remap" in sys.argv:
    # process_equi at the production sizes: 16 frames x 20 views, 512x1024 pano -> 256x256 views (and 1024x2048 -> 512x512)
    import numpy as np
    from imagine360_b200.host import preprocess as P
    for (Hh, Ww, pres) in [(512, 1024, 256), (1024, 2048, 512)]:
        vid = torch.rand(16, 3, Hh, Ww, device="cuda") * 2 - 1
        th = np.linspace(-180, 180, 20)[None]; ph = np.linspace(-60, 60, 20)[None]
        P.process_equi(vid, th, ph, pers_resolution=pres)          # builds + caches the 20 maps on the host
        ms = timeit(lambda: P.process_equi(vid, th, ph, pers_resolution=pres))
        by = vid.numel() * 4 + 2 * vid.numel() + 16 * 20 * 3 * pres * pres * 4 + 20 * pres * pres * 8
        rec(f"process_equi 16x{Hh}x{Ww} -> 20 x {pres}^2", ms, 0, by)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
