"""Launch one kernel shape a few times (for ncu captures).  usage: one_kernel.py gemm M N K [act] | attn imgs N heads hd | conv B H W Ci Co"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagine360_b200 import ops
kind = sys.argv[1]
a = [int(x) for x in sys.argv[2:]]
torch.manual_seed(0)
if kind == "gemm":
    M, N, K = a[:3]; act = a[3] if len(a) > 3 else 0
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N // 2 if act == 1 else N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, w, act=act, out=out)
elif kind == "gemmr":
    M, N, K = a[:3]
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    b = torch.randn(N, device="cuda").bfloat16(); r = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, w, bias=b, resid=r, out=out)
elif kind == "attn":
    imgs, N, heads, hd = a
    C = heads * hd
    qkv = torch.randn(imgs * N, 3 * C, device="cuda").bfloat16(); out = torch.empty(imgs * N, C, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.attention(ops.seq_view(qkv, imgs, N, 0), ops.seq_view(qkv, imgs, N, C), ops.seq_view(qkv, imgs, N, 2 * C),
                               ops.seq_view(out, imgs, N), heads, hd, imgs)
elif kind == "conv":
    B, H, W, Ci, Co = a
    x = torch.randn(B, H, W, Ci, device="cuda").bfloat16(); wp = ops.pack_conv3x3(torch.randn(Co, Ci, 3, 3, device="cuda").bfloat16())
    fn = lambda: ops.conv3x3(x, wp)
if kind == "xattn":
    imgs, N, heads, hd, nk = a; C = heads * hd; fr = 16
    q = torch.randn(imgs * N, C, device="cuda").bfloat16(); kv = torch.randn((imgs // fr) * nk, 2 * C, device="cuda").bfloat16()
    out = torch.zeros(imgs * N, C, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.attention(ops.seq_view(q, imgs, N), ops.seq_view(kv, imgs // fr, nk, 0, share_div=fr),
                               ops.seq_view(kv, imgs // fr, nk, C, share_div=fr), ops.seq_view(out, imgs, N), heads, hd, imgs)
if kind == "temporal":
    B, Fr, D, heads, hd = a
    C = heads * hd
    qkv = torch.randn(B * Fr * D, 3 * C, device="cuda").bfloat16(); out = torch.empty(B * Fr * D, C, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, B, Fr, D, heads, hd)
if kind == "warp":
    b, m, hw, EN, heads = a; hd = 32; C = heads * hd
    pers_kv = torch.randn(b * m * hw, 2 * C, device="cuda").bfloat16(); equi = torch.randn(b * EN, C, device="cuda").bfloat16()
    bias = torch.rand(EN, m * hw, device="cuda").bfloat16() * 2 - 1; out = torch.empty_like(equi)
    fn = lambda: ops.attention(ops.seq_view(equi, b, EN), ops.multiview_view(pers_kv, 1, m, b, hw, 0), ops.multiview_view(pers_kv, 1, m, b, hw, C),
                               ops.seq_view(out, b, EN), heads, hd, b, bias=bias)
if kind == "ln":
    M, C = a
    x = torch.randn(M, C, device="cuda").bfloat16(); g = torch.ones(C, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda").bfloat16()
    fn = lambda: ops.layernorm(x, g, b)
if kind == "gemmln":      # LayerNorm-folded consumer (the producer GEMM runs once, before the loop)
    M, N, K = a[:3]; act = a[3] if len(a) > 3 else 0
    a0 = torch.randn(M, 320, device="cuda").bfloat16(); w0 = torch.randn(K, 320, device="cuda").bfloat16()
    x, st = ops.gemm(a0, w0, rowstats=True)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16(); bb = torch.randn(N, device="cuda").bfloat16() if act else None
    wf, u, c = ops.fold_layernorm(w, bb, torch.ones(K, device="cuda").bfloat16(), torch.zeros(K, device="cuda").bfloat16(), geglu=bool(act))
    out = torch.empty(M, N // 2 if act == 1 else N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm_ln(x, st, wf, u, c, 1e-5, act=act, out=out)
for _ in range(4):
    fn()
torch.cuda.synchronize()
