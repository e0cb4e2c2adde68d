"""A/B of the LayerNorm fold per shape (CUDA events, L2 flushed): producer GEMM with / without row statistics, consumer
LayerNorm + GEMM (two kernels) against the folded GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagine360_b200 import ops
from tools.microbench import timeit  # noqa

BF = torch.bfloat16
def rn(*s): return torch.randn(*s, device="cuda").to(BF)
print("== producers (bias + residual)")
for (M, N, K) in [(655360, 320, 320), (655360, 320, 1280), (163840, 640, 640), (163840, 640, 2560), (40960, 1280, 1280), (40960, 1280, 5120)]:
    a, w, b, r = rn(M, K), rn(N, K), rn(N), rn(M, N)
    out = torch.empty(M, N, device="cuda", dtype=BF)
    t0 = timeit(lambda: ops.gemm(a, w, bias=b, resid=r, out=out))
    t1 = timeit(lambda: ops.gemm(a, w, bias=b, resid=r, out=out, rowstats=True))
    print(f"M={M} N={N} K={K}: plain {t0:.3f} ms  +stats {t1:.3f} ms  ({2.0*M*N*K/t0/1e9:.0f} -> {2.0*M*N*K/t1/1e9:.0f} TF/s)", flush=True)
    del a, w, b, r, out
print("== consumers")
for (M, N, K, act) in [(655360, 960, 320, 0), (655360, 320, 320, 0), (655360, 2560, 320, 1), (163840, 1920, 640, 0), (163840, 5120, 640, 1),
                       (40960, 3840, 1280, 0), (40960, 10240, 1280, 1)]:
    a0, w0, r0 = rn(M, 320), rn(K, 320), rn(M, K)
    x, st = ops.gemm(a0, w0, resid=r0, rowstats=True)
    del a0, w0, r0
    w = rn(N, K) / K ** 0.5
    b = rn(N) if act else None
    gamma, beta = rn(K), rn(K)
    wf, u, c = ops.fold_layernorm(w, b, gamma, beta, geglu=bool(act))
    if act:
        wp, bp = ops.pack_geglu(w, b)
    else:
        wp, bp = w, b
    nrm = torch.empty_like(x)
    out = torch.empty(M, N // 2 if act else N, device="cuda", dtype=BF)
    t_ln = timeit(lambda: ops.layernorm(x, gamma, beta, 1e-5, out=nrm))
    t_g = timeit(lambda: ops.gemm(nrm, wp, bias=bp, act=act, out=out))
    t_f = timeit(lambda: ops.gemm_ln(x, st, wf, u, c, 1e-5, act=act, out=out))
    print(f"M={M} N={N} K={K} act={act}: LN {t_ln:.3f} + GEMM {t_g:.3f} = {t_ln+t_g:.3f} ms | folded {t_f:.3f} ms", flush=True)
    del x, st, w, wf, u, c, nrm, out
