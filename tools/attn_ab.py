"""Self-attention / WarpAttn kernel timings (CUDA events, L2 flushed, median of 10) next to torch SDPA on the same box.
usage: attn_ab.py [warp]      env I360_ATTN_V2=0 selects the one-tile-per-CTA kernel"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from imagine360_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


tag = f"v2={os.environ.get('I360_ATTN_V2', '1')} lib={os.path.basename(os.environ.get('I360_LIB_PATH', 'default'))}"
if "warp" not in sys.argv:
    shapes = [(32, 8192, 5), (640, 1024, 5), (32, 2048, 10), (640, 256, 10), (32, 512, 20), (48, 18432, 5)]
    if "quick" in sys.argv:
        shapes = shapes[:2]
    for imgs, N, heads in shapes:
        hd = 64
        C = heads * hd
        qkv = torch.randn(imgs * N, 3 * C, device="cuda").bfloat16()
        out = torch.empty(imgs * N, C, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.attention(ops.seq_view(qkv, imgs, N, 0), ops.seq_view(qkv, imgs, N, C), ops.seq_view(qkv, imgs, N, 2 * C),
                                   ops.seq_view(out, imgs, N), heads, hd, imgs)
        ms = timeit(fn)
        fl = 4.0 * imgs * heads * N * N * hd
        q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(imgs, N, heads, hd).transpose(1, 2) for i in range(3))
        ms_t = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
        ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(imgs * N, C)
        err = (out.float() - ref.float()).abs().max().item()
        print(f"[{tag}] self-attn imgs={imgs} N={N} heads={heads}: native {ms:.3f} ms ({fl / ms / 1e9:.0f} TFLOP/s) | torch SDPA {ms_t:.3f} ms | max|diff| {err:.4f}", flush=True)
        del qkv, out, q, k, v, ref
else:
    for b, m, hw, EN, heads in [(32, 20, 256, 2048, 20), (32, 20, 256, 2048, 10), (32, 20, 64, 512, 40)]:
        hd = 32
        C = heads * hd
        pers_kv = torch.randn(b * m * hw, 2 * C, device="cuda").bfloat16()
        equi = torch.randn(b * EN, C, device="cuda").bfloat16()
        bias = torch.rand(EN, m * hw, device="cuda").bfloat16() * 2 - 1
        out = torch.empty_like(equi)
        fn = lambda: ops.attention(ops.seq_view(equi, b, EN), ops.multiview_view(pers_kv, 1, m, b, hw, 0), ops.multiview_view(pers_kv, 1, m, b, hw, C),
                                   ops.seq_view(out, b, EN), heads, hd, b, bias=bias)
        ms = timeit(fn)
        fl = 4.0 * b * heads * EN * m * hw * hd
        print(f"[{tag}] warp equi<-pers b={b} heads={heads} Nq={EN} Nk={m * hw}: {ms:.3f} ms ({fl / ms / 1e9:.0f} TFLOP/s)", flush=True)
