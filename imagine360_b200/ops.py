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


def on_device(device):
    """Context manager making ``device`` the current CUDA device (a no-op for CPU devices, so host-only code paths
    and their tests run without a GPU)."""
    import contextlib
    device = torch.device(device)
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


def _chk_dev(t):
    """Kernels are launched on the CURRENT device's current stream: a tensor living elsewhere would be an illegal
    address (or a silent peer access).  The host mirror enters ``torch.cuda.device(tensor.device)`` at its entry
    points (MultiViewBaseModel.forward, AnimationPipeline, AutoencoderKL); direct callers must do the same."""
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"imagine360_b200: tensor on {t.device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device(tensor.device)")


def _chk_bf16(*ts):
    first = True
    for t in ts:
        if t is not None:
            if t.dtype != BF16 or not t.is_cuda:
                raise TypeError(f"expected a CUDA bf16 tensor, got {t.dtype} on {t.device}")
            if first:
                _chk_dev(t)
                first = False


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


class RowStats:
    """Partial (sum, sum of squares) per row of a token matrix, written by the GEMM that produced it (``gemm(...,
    rowstats=True)``) and consumed by ``gemm_ln``: fp32 [slots, M, 2]."""
    __slots__ = ("buf", "slots", "of")

    def __init__(self, buf, slots, of):
        self.buf, self.slots, self.of = buf, slots, of


def gemm(a: torch.Tensor, w: torch.Tensor, bias=None, resid=None, rowvec=None, rowvec_div: int = 1,
         act: int = ACT_NONE, out_scale: float = 1.0, out: torch.Tensor | None = None, rowstats: bool = False):
    """out[M, N] = epilogue(a[M, K] @ w[N, K]^T).  a/out may be row-strided 2-D views.
    ``rowstats=True`` (plain / bias / residual epilogues) returns ``(out, RowStats)``."""
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
    if rowstats:
        assert act == ACT_NONE and rowvec is None and out_scale == 1.0
        slots = int(lib().i360_gemm_rowstats_slots(c_int(M), c_int(N), c_int(K), c_int(int(resid is not None))))
        st = torch.empty((slots, M, 2), dtype=torch.float32, device=a.device)
        rc = lib().i360_gemm_rowstats_bf16(
            _p(a), c_longlong(a.stride(0)), _p(w), c_longlong(w.stride(0)), _p(out), c_longlong(out.stride(0)),
            c_int(M), c_int(N), c_int(K), _p(bias), _p(resid), c_longlong(resid.stride(0) if resid is not None else 0),
            _p(st), _stream())
        check(rc, "i360_gemm_rowstats_bf16")
        return out, RowStats(st, slots, out)
    rc = lib().i360_gemm_bf16(
        _p(a), c_longlong(a.stride(0)), _p(w), c_longlong(w.stride(0)), _p(out), c_longlong(out.stride(0)),
        c_int(M), c_int(N), c_int(K), _p(bias), _p(resid),
        c_longlong(resid.stride(0) if resid is not None else 0),
        _p(rowvec), c_int(rowvec_div), c_int(rowvec.stride(0) if rowvec is not None else 0),
        c_int(act), c_float(out_scale), _stream())
    check(rc, "i360_gemm_bf16")
    return out


def gemm_ln_supported(N: int, K: int, act: int = ACT_NONE) -> bool:
    return int(lib().i360_gemm_ln_supported(c_int(N), c_int(K), c_int(act))) != 0


def fold_layernorm(weight: torch.Tensor, bias, gamma: torch.Tensor, beta: torch.Tensor, geglu: bool = False):
    """(W [N, K], bias [N] | None, LayerNorm gamma / beta [K]) -> (Wf bf16 [N, K], u fp32 [N], c fp32 [N]) for
    :func:`gemm_ln`:  LN(x) W^T + b = rstd * (x Wf^T - mean * u) + c  with Wf = W * gamma, u = rowsum(Wf) taken from the
    bf16-ROUNDED Wf (so that the mean term cancels exactly what the tensor core accumulated), c = W beta + b."""
    w32 = weight.float()
    wf = (w32 * gamma.float()[None, :]).to(BF16)
    u = wf.float().sum(dim=1)
    c = w32 @ beta.float()
    if bias is not None:
        c = c + bias.float()
    if geglu:
        n_total = weight.shape[0]
        half, inner = geglu_block(n_total) // 2, n_total // 2
        idx = torch.cat([torch.cat([torch.arange(t * half, (t + 1) * half), inner + torch.arange(t * half, (t + 1) * half)])
                         for t in range(inner // half)]).to(weight.device)
        wf, u, c = wf.index_select(0, idx), u.index_select(0, idx), c.index_select(0, idx)
    return wf.contiguous(), u.contiguous(), c.contiguous()


def gemm_ln(a: torch.Tensor, stats: RowStats, wf: torch.Tensor, u: torch.Tensor, c: torch.Tensor, eps: float = 1e-5, rowvec=None,
            rowvec_div: int = 1, rowvec_mod: int = 0, act: int = ACT_NONE, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[M, N] = act(LayerNorm(a)[M, K] @ W^T + bias (+ rowvec[(r // rowvec_div) % rowvec_mod])) with the LayerNorm folded
    into the GEMM (see :func:`fold_layernorm`); ``a`` is the UN-normalised token matrix and ``stats`` the row statistics
    its producer wrote (``gemm(..., rowstats=True)``)."""
    assert stats.of is a or (stats.of.data_ptr() == a.data_ptr() and stats.of.shape == a.shape), "row statistics of another tensor"
    assert stats.buf.shape[1] == a.shape[0]
    _chk_bf16(a, wf, out)
    assert a.dim() == 2 and wf.dim() == 2 and a.shape[1] == wf.shape[1] and a.stride(1) == 1 and wf.stride(1) == 1
    assert u.dtype == torch.float32 and c.dtype == torch.float32 and u.is_contiguous() and c.is_contiguous()
    M, K = a.shape
    N = wf.shape[0]
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=BF16, device=a.device)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    if rowvec is not None:
        assert rowvec.dtype == torch.float32 and rowvec.stride(-1) == 1 and rowvec.shape[-1] == N
    rc = lib().i360_gemm_ln_bf16(_p(a), c_longlong(a.stride(0)), _p(wf), c_longlong(wf.stride(0)), _p(out),
                                 c_longlong(out.stride(0)), c_int(M), c_int(N), c_int(K), _p(u), _p(c), c_float(eps),
                                 _p(stats.buf), c_int(stats.slots), _p(rowvec), c_int(rowvec_div), c_int(rowvec_mod),
                                 c_int(rowvec.stride(0) if rowvec is not None else 0), c_int(act), _stream())
    check(rc, "i360_gemm_ln_bf16")
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
            rowvec=None, rowvec_div: int = 1, crop: int = 0, out_scale: float = 1.0, gn_groups: int | None = None,
            chan_stats: bool = False):
    """3x3 / stride 1 / zero-pad 1 convolution on NHWC ``x`` [B, H, W, Cin] -> [B, H, W-2*crop, Cout].

    ``x2``/``x3`` are optional NHWC sources [B, H, W-2*crop, C] (no halo) of a fused 1x1 convolution whose
    weights are the trailing columns of ``w_packed``; ``rowvec`` [B // rowvec_div, Cout] fp32 is added per image (temb)."""
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
    if chan_stats:
        # the epilogue also accumulates PER-CHANNEL statistics of the output -> (out, stats [B, Cout, 2] fp64); any group size
        st = torch.empty((B, Cout, 2), dtype=torch.float64, device=x.device)
        rc = lib().i360_conv3x3_chanstats_bf16(
            _p(x), c_int(B), c_int(H), c_int(W), c_int(Cin), _p(x2), c_int(C2), _p(x3), c_int(C3),
            _p(w_packed), c_int(Cout), _p(out), c_int(crop), _p(bias), _p(resid), _p(rowvec),
            c_int(rowvec_div), c_int(rowvec.shape[-1] if rowvec is not None else 0), c_float(out_scale), _p(st), _stream())
        check(rc, "i360_conv3x3_chanstats_bf16")
        return out, st
    if gn_groups is not None:
        # the epilogue also accumulates the GroupNorm statistics of the output -> (out, stats [B, 32, 2] fp64)
        assert x3 is None and rowvec is None and crop == 0 and out_scale == 1.0
        stats = torch.empty((B, gn_groups, 2), dtype=torch.float64, device=x.device)
        rc = lib().i360_conv3x3_gnstats_bf16(_p(x), c_int(B), c_int(H), c_int(W), c_int(Cin), _p(x2), c_int(C2), _p(w_packed),
                                             c_int(Cout), _p(out), _p(bias), _p(resid), c_int(gn_groups), _p(stats), _stream())
        check(rc, "i360_conv3x3_gnstats_bf16")
        return out, stats
    rc = lib().i360_conv3x3_bf16(
        _p(x), c_int(B), c_int(H), c_int(W), c_int(Cin), _p(x2), c_int(C2), _p(x3), c_int(C3),
        _p(w_packed), c_int(Cout), _p(out), c_int(crop), _p(bias), _p(resid), _p(rowvec),
        c_int(rowvec_div), c_int(rowvec.shape[-1] if rowvec is not None else 0),
        c_float(out_scale), _stream())
    check(rc, "i360_conv3x3_bf16")
    return out


def conv3x3_s2(x: torch.Tensor, w_packed: torch.Tensor, bias=None, pad_lo: int = 1, crop: int = 0) -> torch.Tensor:
    """3x3 / stride 2 conv of NHWC ``x`` [B, H, W, Cin] -> [B, H/2, W/2 - 2 crop, Cout] without an im2col buffer (TMA boxes
    with traversal stride 2); ``pad_lo`` = 1: zero pad 1 all round, 0: pad (0, 1) like the VAE encoder."""
    _chk_bf16(x, w_packed, bias)
    assert x.dim() == 4 and x.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == 9 * Cin and w_packed.is_contiguous() and H % 2 == 0 and W % 2 == 0
    out = torch.empty((B, H // 2, W // 2 - 2 * crop, Cout), dtype=BF16, device=x.device)
    rc = lib().i360_conv3x3_s2_bf16(_p(x), c_int(B), c_int(H), c_int(W), c_int(Cin), _p(w_packed), c_int(Cout), _p(out),
                                    c_int(pad_lo), c_int(crop), _p(bias), _stream())
    check(rc, "i360_conv3x3_s2_bf16")
    return out


def pack_upsample_conv(weight: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [4, Cout, 4 * Cin] bf16: the pre-summed 2x2-tap weights of the four output parities of
    "nearest x2 upsample -> conv3x3" (see i360_conv_upsample2x_bf16); sums in fp32, one rounding."""
    w = weight.detach().float()
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}          # parity -> (taps on input row i + a - 1, taps on input row i + a)
    out = []
    for a in (0, 1):
        for b in (0, 1):
            taps = []
            for dr in (0, 1):
                for dc in (0, 1):
                    t = sum(w[:, :, kh, kw] for kh in rows[a][dr] for kw in rows[b][dc])      # [Cout, Cin]
                    taps.append(t)
            out.append(torch.cat(taps, dim=1))
    return torch.stack(out).to(BF16).contiguous()


def conv_upsample2x(x: torch.Tensor, w_eff: torch.Tensor, bias=None, crop: int = 0, gn_groups: int | None = None):
    """nearest x2 upsample + 3x3 conv of NHWC ``x`` [B, H, W, Cin] (``crop`` circular halo columns per side included in W)
    -> [B, 2H, 2(W - 2 crop), Cout], in sub-pixel form (no upsampled tensor, 4/9 of the FLOPs)."""
    _chk_bf16(x, w_eff, bias)
    assert x.dim() == 4 and x.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w_eff.shape[1]
    assert w_eff.shape == (4, Cout, 4 * Cin) and w_eff.is_contiguous()
    out = torch.empty((B, 2 * H, 2 * (W - 2 * crop), Cout), dtype=BF16, device=x.device)
    if gn_groups is not None:
        assert crop == 0
        stats = torch.empty((B, gn_groups, 2), dtype=torch.float64, device=x.device)
        rc = lib().i360_conv_upsample2x_gnstats_bf16(_p(x), c_int(B), c_int(H), c_int(W), c_int(Cin), _p(w_eff), c_int(Cout), _p(out),
                                                     _p(bias), c_int(gn_groups), _p(stats), _stream())
        check(rc, "i360_conv_upsample2x_gnstats_bf16")
        return out, stats
    rc = lib().i360_conv_upsample2x_bf16(_p(x), c_int(B), c_int(H), c_int(W), c_int(Cin), _p(w_eff), c_int(Cout), _p(out),
                                         c_int(crop), _p(bias), _stream())
    check(rc, "i360_conv_upsample2x_bf16")
    return out


def conv3x3_uses_halo(B: int, H: int, W: int, Cin: int, resid: bool = False, rowvec: bool = False, extra: bool = False) -> bool:
    return bool(lib().i360_conv3x3_uses_halo(c_int(B), c_int(H), c_int(W), c_int(Cin), c_int(int(resid)), c_int(int(rowvec)),
                                             c_int(int(extra))))


def conv3x3_halo_policy(on: int = -1, tol: float = -1.0, allow_extra: int = -1, min_hw: int = -1) -> None:
    """Selection rule of the halo conv kernels (tests / A-B runs); negative = keep."""
    f = lib().i360_conv3x3_halo_policy
    f.restype = None
    f(c_int(on), ctypes.c_double(tol), c_int(allow_extra), c_int(min_hw))


# ------------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------------
def groupnorm(x1: torch.Tensor, gamma, beta, groups: int, eps: float, silu: bool, x2=None, pad: int = 0,
              stats_pad: int | None = None, stats: torch.Tensor | None = None, chan_stats: torch.Tensor | None = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over the channel concat of NHWC ``x1`` (and ``x2``), circularly padded by ``pad`` columns.
    Statistics are per image over the padded tensor (``stats_pad`` overrides the pad used for statistics); ``stats``
    [B, groups, 2] fp64 from the producing conv's epilogue (``conv3x3(..., gn_groups=)``) skips the statistics pass."""
    _chk_bf16(x1, x2, gamma, beta)
    assert x1.dim() == 4 and x1.is_contiguous() and (x2 is None or (x2.is_contiguous() and x2.shape[:3] == x1.shape[:3]))
    B, H, W, C1 = x1.shape
    C2 = x2.shape[-1] if x2 is not None else 0
    C = C1 + C2
    sp = pad if stats_pad is None else stats_pad
    if chan_stats is not None:
        # per-channel statistics [B, C, 2] fp64 from the producing conv's epilogue (conv3x3(..., chan_stats=True))
        assert x2 is None and pad == 0 and sp == 0 and chan_stats.shape == (B, C, 2) and chan_stats.dtype == torch.float64
        out = torch.empty((B, H, W, C), dtype=BF16, device=x1.device)
        rc = lib().i360_groupnorm_apply_chanstats(_p(x1), c_int(C), c_int(B), c_int(H), c_int(W), c_int(groups), _p(chan_stats),
                                                  _p(gamma), _p(beta), c_float(eps), c_int(1 if silu else 0), _p(out), _stream())
        check(rc, "i360_groupnorm_apply_chanstats")
        return out
    if stats is None:
        stats = torch.empty((B, groups, 2), dtype=torch.float64, device=x1.device)
        rc = lib().i360_groupnorm_stats(_p(x1), c_int(C1), _p(x2), c_int(C2), c_int(B), c_int(H), c_int(W), c_int(sp),
                                        c_int(groups), _p(stats), _stream())
        check(rc, "i360_groupnorm_stats")
    else:
        assert x2 is None and pad == 0 and sp == 0 and stats.shape == (B, groups, 2) and stats.dtype == torch.float64
    out = torch.empty((B, H, W + 2 * pad, C), dtype=BF16, device=x1.device)
    count = float(H * (W + 2 * sp) * (C // groups)) if sp != pad else 0.0
    rc = lib().i360_groupnorm_apply(_p(x1), c_int(C1), _p(x2), c_int(C2), c_int(B), c_int(H), c_int(W), c_int(pad),
                                    c_int(groups), _p(stats), ctypes.c_double(count), _p(gamma), _p(beta),
                                    c_float(eps), c_int(1 if silu else 0), _p(out), _stream())
    check(rc, "i360_groupnorm_apply")
    return out


def layernorm(x: torch.Tensor, gamma, beta, eps: float = 1e-5, pre_add=None, pre_index=(1, 1, 0, 1), post_add=None,
              post_div: int = 1, post_mod: int = 1, out=None) -> torch.Tensor:
    """LayerNorm over the last dim of 2-D ``x`` [M, C] (row-strided views allowed).
    ``pre_add`` bf16 [*, C] is added first at row ((r // a) % b) * c + r % d with (a, b, c, d) = pre_index;
    ``post_add`` fp32 [post_mod, C] is added last at row (r // post_div) % post_mod."""
    _chk_bf16(x, gamma, beta, pre_add, out)
    assert x.dim() == 2 and x.stride(1) == 1
    M, C = x.shape
    if out is None:
        out = torch.empty((M, C), dtype=BF16, device=x.device)
    if post_add is not None:
        assert post_add.dtype == torch.float32 and post_add.is_contiguous() and post_add.shape[-1] == C
    if pre_add is not None:
        assert pre_add.is_contiguous() and pre_add.shape[-1] == C
    a, b, c, d = pre_index
    rc = lib().i360_layernorm(_p(x), c_longlong(x.stride(0)), _p(out), c_longlong(out.stride(0)), c_longlong(M), c_int(C),
                              _p(gamma), _p(beta), c_float(eps), _p(pre_add), c_int(a), c_int(b), c_int(c), c_int(d),
                              _p(post_add), c_int(post_div), c_int(post_mod), _stream())
    check(rc, "i360_layernorm")
    return out


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
class TokenView(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("channels", c_int), ("col0", c_int), ("d1", c_int), ("d2", c_int), ("d3", c_int),
                ("s1", c_longlong), ("s2", c_longlong), ("s3", c_longlong), ("A", c_int), ("Bdiv", c_int),
                ("mul", c_int), ("ext3", c_int)]


def seq_view(t: torch.Tensor, n_seq: int, n_tok: int, col0: int = 0, share_div: int = 1) -> TokenView:
    """Rows of 2-D ``t`` are ``n_seq`` consecutive sequences of ``n_tok`` tokens; batch item bi uses sequence
    bi // share_div (share_div = frames when K/V are shared by all frames of a clip)."""
    assert t.dim() == 2 and t.stride(1) == 1
    ld = t.stride(0)
    v = TokenView(t.data_ptr(), t.shape[1], col0, n_tok, n_seq, 1, ld, ld * n_tok, ld * n_tok * n_seq, 0, share_div,
                  0, 1)
    v.src = ("seq", t, n_seq, n_tok, col0, share_div)       # python-side description (imagine360_b200.debug reads it)
    return v


def multiview_view(t: torch.Tensor, n_clip: int, n_view: int, n_frame: int, n_tok: int, col0: int = 0) -> TokenView:
    """Rows of ``t`` are ordered (clip, view, frame, token); batch item bi = clip * n_frame + frame attends over
    the (view, token) axis of its frame (WarpAttn's '(b m) c f h w -> (b f) (m h w) c')."""
    assert t.dim() == 2 and t.stride(1) == 1
    ld = t.stride(0)
    v = TokenView(t.data_ptr(), t.shape[1], col0, n_tok, n_frame, n_clip * n_view, ld, ld * n_tok,
                  ld * n_tok * n_frame, n_frame, 1, n_view, n_view)
    v.src = ("multiview", t, n_clip, n_view, n_frame, n_tok, col0)
    return v


def _tile_rule(d1: int, ext3: int):
    """(box1, box3, n1) exactly as attention.cu::fill_operand chooses them."""
    if ext3 > 1 and d1 < 128 and 128 % d1 == 0:
        return d1, 128 // d1, 1
    return 128, 1, (d1 + 127) // 128


def tile_attention_bias(bias: torch.Tensor, q: TokenView, k: TokenView) -> torch.Tensor:
    """Dense [Nq, Nk] bias (sequence order (view, token)) -> the kernel's tile layout.  A no-op whenever every view
    is a multiple of 128 tokens or divides 128 with the views filling whole tiles."""
    def axis(t, dim, d1, ext3):
        b1, b3, n1 = _tile_rule(d1, ext3)
        if b3 > 1 or d1 % 128 == 0:
            return t                                   # already contiguous in tile order
        shp = list(t.shape)
        t = t.unflatten(dim, (ext3, d1))
        pad = [0, 0] * (t.dim() - dim - 2) + [0, n1 * 128 - d1]
        t = F.pad(t, pad)
        return t.flatten(dim, dim + 1)
    import torch.nn.functional as F
    out = axis(bias, 0, q.d1, max(q.ext3, 1))
    out = axis(out, 1, k.d1, max(k.ext3, 1))
    if out.shape[1] % 8:
        out = F.pad(out, (0, 8 - out.shape[1] % 8))
    return out.contiguous()


def attention(q: TokenView, k: TokenView, v: TokenView, o: TokenView, heads: int, head_dim: int, batch: int,
              scale: float | None = None, bias: torch.Tensor | None = None, accumulate: bool = False) -> None:
    if bias is not None:
        _chk_bf16(bias)
        assert bias.dim() == 2
        bias = tile_attention_bias(bias, q, k)       # identity for every production shape
    rc = lib().i360_attention_bf16(ctypes.byref(q), ctypes.byref(k), ctypes.byref(v), ctypes.byref(o), c_int(heads),
                                   c_int(head_dim), c_int(batch), c_float(scale if scale is not None else head_dim ** -0.5),
                                   _p(bias), c_int(bias.shape[0] if bias is not None else 0),
                                   c_int(bias.shape[1] if bias is not None else 0), c_int(1 if accumulate else 0), _stream())
    check(rc, "i360_attention_bf16")


def attention_item_bias(q: TokenView, k: TokenView, v: TokenView, o: TokenView, heads: int, head_dim: int, batch: int,
                        bias: torch.Tensor, scale: float | None = None) -> None:
    """Attention with one [Nq, Nk_padded] bias block per (batch item, head): ``bias`` is [batch * heads, Nq, ldb] bf16."""
    _chk_bf16(bias)
    assert bias.dim() == 3 and bias.is_contiguous() and bias.shape[0] == batch * heads and bias.shape[1] == q.d1
    rc = lib().i360_attention_item_bias_bf16(ctypes.byref(q), ctypes.byref(k), ctypes.byref(v), ctypes.byref(o), c_int(heads),
                                             c_int(head_dim), c_int(batch),
                                             c_float(scale if scale is not None else head_dim ** -0.5), _p(bias),
                                             c_int(bias.shape[1]), c_int(bias.shape[2]), _stream())
    check(rc, "i360_attention_item_bias_bf16")


def relpos_bias(qkv: torch.Tensor, col0: int, items: int, heads: int, head_dim: int, S: int, rel_h: torch.Tensor,
                rel_w: torch.Tensor) -> torch.Tensor:
    """SAM's decomposed relative position bias for ``items`` sequences of S x S tokens -> [items * heads, S*S, ldb] bf16."""
    _chk_bf16(qkv, rel_h, rel_w)
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[0] == items * S * S
    assert rel_h.shape == (2 * S - 1, head_dim) and rel_w.shape == (2 * S - 1, head_dim) and rel_h.is_contiguous() and rel_w.is_contiguous()
    ldb = -(-S * S // 8) * 8
    bias = torch.empty((items * heads, S * S, ldb), dtype=BF16, device=qkv.device)
    rc = lib().i360_relpos_bias_bf16(_p(qkv), c_longlong(qkv.stride(0)), c_int(col0), c_int(items), c_int(heads), c_int(head_dim),
                                     c_int(S), _p(rel_h), _p(rel_w), _p(bias), c_int(ldb), _stream())
    check(rc, "i360_relpos_bias_bf16")
    return bias


def cross_attention_text_ip_supported(head_dim: int, nt: int, ni: int) -> bool:
    ntp, nip = -(-nt // 16) * 16, -(-ni // 16) * 16
    return head_dim == 64 and ntp <= 96 and nip <= 96 and -(-ntp // 64) + -(-nip // 64) <= 3


def cross_attention_text_ip(q: torch.Tensor, out: torch.Tensor, kv_text: torch.Tensor, nt: int, kv_ip: torch.Tensor, ni: int,
                            n_ctx: int, heads: int, head_dim: int, scale: float | None = None) -> None:
    """``out = softmax(q Kt^T) Vt + softmax(q Ki^T) Vi`` per head; rows of ``q`` are ``n_ctx`` clip elements times
    (frames x tokens); ``kv_*`` hold [K | V] of each element's text / image-prompt tokens (attention.py:119-148)."""
    _chk_bf16(q, out, kv_text, kv_ip)
    assert q.dim() == 2 and q.stride(1) == 1 and out.stride(1) == 1 and kv_text.stride(1) == 1 and kv_ip.stride(1) == 1
    assert q.shape[1] == heads * head_dim and kv_text.shape == (n_ctx * nt, 2 * heads * head_dim)
    assert kv_ip.shape == (n_ctx * ni, 2 * heads * head_dim) and q.shape[0] % n_ctx == 0
    rc = lib().i360_cross_attention_text_ip_bf16(_p(q), c_longlong(q.stride(0)), _p(out), c_longlong(out.stride(0)),
                                                 c_longlong(q.shape[0]), c_int(n_ctx), _p(kv_text), c_longlong(kv_text.stride(0)),
                                                 c_int(nt), _p(kv_ip), c_longlong(kv_ip.stride(0)), c_int(ni), c_int(heads),
                                                 c_int(head_dim), c_float(scale if scale is not None else head_dim ** -0.5),
                                                 _stream())
    check(rc, "i360_cross_attention_text_ip_bf16")


def temporal_attention(q, k, v, out, B: int, F: int, D: int, heads: int, head_dim: int) -> None:
    """q/k/v/out: 2-D bf16 views with rows ordered (b, f, d) and heads*head_dim columns."""
    _chk_bf16(q, k, v, out)
    rc = lib().i360_temporal_attention_bf16(_p(q), c_longlong(q.stride(0)), _p(k), c_longlong(k.stride(0)), _p(v),
                                            c_longlong(v.stride(0)), _p(out), c_longlong(out.stride(0)), c_int(B),
                                            c_int(F), c_int(D), c_int(heads), c_int(head_dim),
                                            c_float(head_dim ** -0.5), _stream())
    check(rc, "i360_temporal_attention_bf16")


# ------------------------------------------------------------------------------------------------
# data movement / elementwise
# ------------------------------------------------------------------------------------------------
def upsample2x(x: torch.Tensor, pad_in: int = 0) -> torch.Tensor:
    _chk_bf16(x)
    B, H, W, C = x.shape
    out = torch.empty((B, 2 * H, 2 * (W + 2 * pad_in), C), dtype=BF16, device=x.device)
    check(lib().i360_upsample2x_nhwc(_p(x), _p(out), c_int(B), c_int(H), c_int(W), c_int(C), c_int(pad_in), _stream()),
          "i360_upsample2x_nhwc")
    return out


def softmax_rows(x: torch.Tensor) -> torch.Tensor:
    _chk_bf16(x)
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.empty_like(x)
    check(lib().i360_softmax_rows_bf16(_p(x), _p(out), c_longlong(x.stride(0)), c_longlong(x.shape[0]), c_int(x.shape[1]),
                                       _stream()), "i360_softmax_rows_bf16")
    return out


def im2col_s2(x: torch.Tensor, circular: bool, pad_lo: int = 1) -> torch.Tensor:
    _chk_bf16(x)
    B, H, W, C = x.shape
    out = torch.empty((B * (H // 2) * (W // 2), 9 * C), dtype=BF16, device=x.device)
    check(lib().i360_im2col3x3_s2_nhwc(_p(x), _p(out), c_int(B), c_int(H), c_int(W), c_int(C), c_int(1 if circular else 0),
                                       c_int(pad_lo), _stream()), "i360_im2col3x3_s2_nhwc")
    return out


def axpby(x: torch.Tensor, y: torch.Tensor | None, a: float, b: float) -> torch.Tensor:
    _chk_bf16(x, y)
    assert x.is_contiguous() and (y is None or (y.is_contiguous() and y.shape == x.shape))
    out = torch.empty_like(x)
    check(lib().i360_axpby_bf16(_p(x), _p(y), _p(out), c_float(a), c_float(b), c_longlong(x.numel()), _stream()),
          "i360_axpby_bf16")
    return out


def cfg_ddim_step(latent, pred_uncond, pred_cond, guidance: float, sa: float, sb: float, sap: float, sbp: float, out=None):
    """CFG combine + DDIM v-prediction update in one pass; ``out`` may alias ``latent`` (element-wise kernel)."""
    _chk_bf16(latent, pred_uncond, pred_cond, out)
    assert latent.is_contiguous() and pred_uncond.is_contiguous() and pred_cond.is_contiguous()
    assert latent.shape == pred_uncond.shape == pred_cond.shape
    if out is None:
        out = torch.empty_like(latent)
    assert out.shape == latent.shape and out.is_contiguous()
    check(lib().i360_cfg_ddim_step_bf16(_p(latent), _p(pred_uncond), _p(pred_cond), _p(out), c_float(guidance), c_float(sa),
                                        c_float(sb), c_float(sap), c_float(sbp), c_longlong(latent.numel()), _stream()),
          "i360_cfg_ddim_step_bf16")
    return out


def avgpool_frames4(x: torch.Tensor) -> torch.Tensor:
    """[B, F, D, C] -> [B, F // 4, D, C]"""
    _chk_bf16(x)
    assert x.is_contiguous() and x.dim() == 4
    B, F, D, C = x.shape
    out = torch.empty((B, F // 4, D, C), dtype=BF16, device=x.device)
    check(lib().i360_avgpool_frames4_bf16(_p(x), _p(out), c_int(B), c_int(F), c_longlong(D * C), _stream()),
          "i360_avgpool_frames4_bf16")
    return out


def grid_sample(img: torch.Tensor, grid: torch.Tensor, nearest: bool = False) -> torch.Tensor:
    """img [N, C, Hi, Wi] fp32, grid [N, Ho, Wo, 2] fp32 normalised (x, y), align_corners=True, zeros padding."""
    assert img.dtype == torch.float32 and grid.dtype == torch.float32 and img.is_cuda and grid.is_cuda
    _chk_dev(img)
    img, grid = img.contiguous(), grid.contiguous()
    N, C, Hi, Wi = img.shape
    Ho, Wo = grid.shape[1:3]
    out = torch.empty((N, C, Ho, Wo), dtype=torch.float32, device=img.device)
    check(lib().i360_grid_sample_f32(_p(img), _p(grid), _p(out), c_int(N), c_int(C), c_int(Hi), c_int(Wi), c_int(Ho),
                                     c_int(Wo), c_int(1 if nearest else 0), _stream()), "i360_grid_sample_f32")
    return out
