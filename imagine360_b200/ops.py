"""Torch-tensor front door to the C-ABI kernels.

PyTorch is plumbing here (device memory, streams); every function below forwards raw device
pointers to ``libimagine360_b200.so``.  All activations are bf16 and channels-last
(``[images, H, W, C]`` / ``[tokens, C]``).
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_longlong, c_void_p

import torch

from ._lib import check, lib

BF16 = torch.bfloat16


def _p(t):
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_bf16(*ts):
    for t in ts:
        if t is not None:
            if t.dtype != BF16 or not t.is_cuda:
                raise TypeError(f"expected a CUDA bf16 tensor, got {t.dtype} on {t.device}")


ACT_NONE, ACT_GEGLU, ACT_GELU, ACT_SILU = 0, 1, 2, 3


def geglu_block(n_total: int) -> int:
    return int(lib().i360_gemm_geglu_block(c_int(n_total)))


def pack_geglu(weight: torch.Tensor, bias: torch.Tensor | None):
    """Re-order a GEGLU projection [2*inner, K] (values then gates, activations.py:93-122) into
    per-tile [values | gates] blocks so one accumulator tile holds matching halves."""
    n_total = weight.shape[0]
    bn = geglu_block(n_total)
    half = bn // 2
    inner = n_total // 2
    idx = []
    for t in range(inner // half):
        idx.append(torch.arange(t * half, (t + 1) * half))
        idx.append(inner + torch.arange(t * half, (t + 1) * half))
    idx = torch.cat(idx).to(weight.device)
    w = weight.index_select(0, idx).contiguous()
    b = bias.index_select(0, idx).contiguous() if bias is not None else None
    return w, b


def gemm(a: torch.Tensor, w: torch.Tensor, bias=None, resid=None, rowvec=None, rowvec_div: int = 1,
         act: int = ACT_NONE, out_scale: float = 1.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[M, N] = epilogue(a[M, K] @ w[N, K]^T).  a/out may be row-strided 2-D views."""
    _chk_bf16(a, w, bias, resid, out)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=BF16, device=a.device)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    if resid is not None:
        assert resid.shape == (M, n_out) and resid.stride(1) == 1
    if rowvec is not None:
        assert rowvec.dtype == torch.float32 and rowvec.stride(-1) == 1
    rc = lib().i360_gemm_bf16(
        _p(a), c_longlong(a.stride(0)), _p(w), c_longlong(w.stride(0)), _p(out), c_longlong(out.stride(0)),
        c_int(M), c_int(N), c_int(K), _p(bias), _p(resid),
        c_longlong(resid.stride(0) if resid is not None else 0),
        _p(rowvec), c_int(rowvec_div), c_int(rowvec.stride(0) if rowvec is not None else 0),
        c_int(act), c_float(out_scale), _stream())
    check(rc, "i360_gemm_bf16")
    return out


def pack_conv3x3(weight: torch.Tensor, *shortcuts: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] (+ optional 1x1 shortcut weights [Cout, Ci, 1, 1]) -> [Cout, 9*Cin + sum Ci]
    with the 3x3 part ordered (kh, kw, cin) to match the tap loop of the kernel."""
    co = weight.shape[0]
    parts = [weight.permute(0, 2, 3, 1).reshape(co, -1)]
    for s in shortcuts:
        parts.append(s.reshape(co, -1))
    return torch.cat(parts, dim=1).to(BF16).contiguous()


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, bias=None, x2=None, x3=None, resid=None,
            rowvec=None, rowvec_div: int = 1, crop: int = 0, out_scale: float = 1.0) -> torch.Tensor:
    """3x3 / stride 1 / zero-pad 1 convolution on NHWC ``x`` [B, H, W, Cin] -> [B, H, W-2*crop, Cout].

    ``x2``/``x3`` are optional NHWC sources of a fused 1x1 convolution whose weights are the trailing
    columns of ``w_packed``; ``rowvec`` [B // rowvec_div, Cout] fp32 is added per image (temb)."""
    _chk_bf16(x, w_packed, bias, x2, x3, resid)
    assert x.dim() == 4 and x.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    C2 = x2.shape[-1] if x2 is not None else 0
    C3 = x3.shape[-1] if x3 is not None else 0
    assert w_packed.shape[1] == 9 * Cin + C2 + C3 and w_packed.is_contiguous()
    out = torch.empty((B, H, W - 2 * crop, Cout), dtype=BF16, device=x.device)
    if resid is not None:
        assert resid.shape == out.shape and resid.is_contiguous()
    if rowvec is not None:
        assert rowvec.dtype == torch.float32 and rowvec.is_contiguous()
    rc = lib().i360_conv3x3_bf16(
        _p(x), c_int(B), c_int(H), c_int(W), c_int(Cin), _p(x2), c_int(C2), _p(x3), c_int(C3),
        _p(w_packed), c_int(Cout), _p(out), c_int(crop), _p(bias), _p(resid), _p(rowvec),
        c_int(rowvec_div), c_int(rowvec.shape[-1] if rowvec is not None else 0),
        c_float(out_scale), _stream())
    check(rc, "i360_conv3x3_bf16")
    return out
