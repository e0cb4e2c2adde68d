"""Bisecting aid ("env flag to switch reference/native per op"): run chosen ops of the native path
through plain torch-CUDA math instead of the sm_100a kernels, to localise a numerical difference to one kernel family.

    I360_REFERENCE_OPS=attention,groupnorm python ...        # or:  with debug.reference_ops("gemm"): ...

This is a DEBUGGING tool and not a fallback: nothing in the product imports this module, the switch is off unless asked
for, it still needs the GPU, and it is orders of magnitude slower.  Switchable: gemm (with the LayerNorm-folding gemm_ln), conv3x3, groupnorm, layernorm,
attention, temporal_attention, cfg_ddim_step."""
from __future__ import annotations

import contextlib
import os

import torch
import torch.nn.functional as F

from . import ops

BF16 = torch.bfloat16


def _gemm(a, w, bias=None, resid=None, rowvec=None, rowvec_div=1, act=ops.ACT_NONE, out_scale=1.0, out=None, rowstats=False):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if rowvec is not None:
        y = y + rowvec[torch.arange(a.shape[0], device=a.device) // rowvec_div]
    if act == ops.ACT_GEGLU:
        # packed [values | gates] blocks of geglu_block(N) rows (ops.pack_geglu)
        bn = ops.geglu_block(w.shape[0])
        y = y.view(a.shape[0], -1, 2, bn // 2)
        y = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(a.shape[0], -1)
    elif act == ops.ACT_GELU:
        y = F.gelu(y)
    elif act == ops.ACT_SILU:
        y = F.silu(y)
    if resid is not None:
        y = y + resid.float()
    yf = y * out_scale
    y = yf.to(BF16)
    if out is not None:
        out.copy_(y)
        y = out
    if rowstats:        # one slot: (sum, sum of squares) of the fp32 values, like the kernel's epilogue
        return y, ops.RowStats(torch.stack([yf.sum(1), (yf * yf).sum(1)], -1)[None].contiguous(), 1, y)
    return y


def _gemm_ln(a, stats, wf, u, c, eps=1e-5, rowvec=None, rowvec_div=1, rowvec_mod=0, act=ops.ACT_NONE, out=None):
    """The folded LayerNorm -> Linear with the kernel's own formula: rstd * (a Wf^T - mean * u) + c (+ rowvec)."""
    s = stats.buf.sum(0)
    k = a.shape[1]
    mean = s[:, 0] / k
    rstd = torch.rsqrt((s[:, 1] / k - mean * mean).clamp_min(0) + eps)
    y = rstd[:, None] * (a.float() @ wf.float().t() - mean[:, None] * u[None]) + c[None]
    if rowvec is not None:
        r = torch.arange(a.shape[0], device=a.device) // rowvec_div
        y = y + rowvec[r % rowvec_mod if rowvec_mod else r]
    if act == ops.ACT_GEGLU:
        bn = ops.geglu_block(wf.shape[0])
        y = y.view(a.shape[0], -1, 2, bn // 2)
        y = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(a.shape[0], -1)
    y = y.to(BF16)
    if out is not None:
        out.copy_(y)
        return out
    return y


def _conv3x3(x, w_packed, bias=None, x2=None, x3=None, resid=None, rowvec=None, rowvec_div=1, crop=0, out_scale=1.0, gn_groups=None,
             chan_stats=False):
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    w = w_packed[:, : 9 * Cin].float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None, padding=1)
    if crop:
        y = y[..., crop:-crop]
    extra = [t for t in (x2, x3) if t is not None]
    if extra:
        ws = w_packed[:, 9 * Cin:].float()
        y = y + torch.einsum("bhwc,oc->bohw", torch.cat(extra, -1).float(), ws)
    if bias is not None:
        y = y + bias.float()[None, :, None, None]
    if rowvec is not None:
        y = y + rowvec[torch.arange(B, device=x.device) // rowvec_div][:, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if resid is not None:
        y = y + resid.float()
    y = (y * out_scale).to(BF16).contiguous()
    if chan_stats or gn_groups is not None:     # the statistics the epilogue leaves: sums of the stored values
        v = y.double().view(B, -1, Cout) if chan_stats else y.double().view(B, -1, gn_groups, Cout // gn_groups)
        red = (1,) if chan_stats else (1, 3)
        st = torch.stack([v.sum(dim=red), (v * v).sum(dim=red)], -1)
        return y, st
    return y


def _groupnorm(x1, gamma, beta, groups, eps, silu, x2=None, pad=0, stats_pad=None, stats=None, chan_stats=None):
    # ``stats`` / ``chan_stats`` (statistics from the producing conv's epilogue) are recomputed from the tensor here
    x = x1 if x2 is None else torch.cat([x1, x2], -1)
    sp = pad if stats_pad is None else stats_pad

    def padw(t, p):
        return t if p == 0 else torch.cat([t[:, :, -p:], t, t[:, :, :p]], 2)

    xs = padw(x, sp).float()
    B, H, W, C = xs.shape
    g = xs.view(B, H * W, groups, C // groups)
    mean = g.mean(dim=(1, 3), keepdim=True)
    var = g.var(dim=(1, 3), unbiased=False, keepdim=True)
    xo = padw(x, pad).float()
    y = ((xo.view(B, -1, groups, C // groups) - mean) * torch.rsqrt(var + eps)).view(xo.shape) * gamma.float() + beta.float()
    return (F.silu(y) if silu else y).to(BF16).contiguous()


def _layernorm(x, gamma, beta, eps=1e-5, pre_add=None, pre_index=(1, 1, 0, 1), post_add=None, post_div=1, post_mod=1, out=None):
    r = torch.arange(x.shape[0], device=x.device)
    xf = x
    if pre_add is not None:
        a, b, c, d = pre_index
        xf = x + pre_add[((r // a) % b) * c + r % d]
    y = F.layer_norm(xf.float(), (x.shape[1],), gamma.float(), beta.float(), eps)
    if post_add is not None:
        y = y.to(BF16).float() + post_add[(r // post_div) % post_mod]
    y = y.to(BF16)
    if out is not None:
        out.copy_(y)
        return out
    return y


def _gather(view, heads, hd, batch):
    kind = view.src[0]
    if kind == "seq":
        _, t, n_seq, n_tok, col0, share = view.src
        x = t[:, col0:col0 + heads * hd].reshape(n_seq, n_tok, heads, hd)
        return x[torch.arange(batch, device=t.device) // share]
    _, t, n_clip, n_view, n_frame, n_tok, col0 = view.src
    x = t[:, col0:col0 + heads * hd].reshape(n_clip, n_view, n_frame, n_tok, heads, hd)
    return x.permute(0, 2, 1, 3, 4, 5).reshape(n_clip * n_frame, n_view * n_tok, heads, hd)


def _attention(q, k, v, o, heads, head_dim, batch, scale=None, bias=None, accumulate=False):
    qh, kh, vh = (_gather(t, heads, head_dim, batch).float().transpose(1, 2) for t in (q, k, v))
    y = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=None if bias is None else bias.float(), scale=scale)
    y = y.transpose(1, 2).to(BF16)                                    # [batch, Nq, heads, hd]
    if o.src[0] == "seq":
        _, t, n_seq, n_tok, col0, _ = o.src
        dst = t[:, col0:col0 + heads * head_dim].view(n_seq, n_tok, heads * head_dim)
        y = y.reshape(batch, n_tok, -1)
    else:
        _, t, n_clip, n_view, n_frame, n_tok, col0 = o.src
        dst = t[:, col0:col0 + heads * head_dim].view(n_clip, n_view, n_frame, n_tok, heads * head_dim)
        y = y.reshape(n_clip, n_frame, n_view, n_tok, -1).permute(0, 2, 1, 3, 4)
    dst.copy_(dst + y if accumulate else y)


def _temporal_attention(q, k, v, out, B, Fr, D, heads, head_dim):
    def bd(t):
        return t.reshape(B, Fr, D, heads, head_dim).permute(0, 2, 3, 1, 4).float()      # [B, D, heads, F, hd]
    y = F.scaled_dot_product_attention(bd(q), bd(k), bd(v))
    out.copy_(y.permute(0, 3, 1, 2, 4).reshape(B * Fr * D, heads * head_dim).to(BF16))


def _cfg_ddim_step(latent, pred_uncond, pred_cond, guidance, sa, sb, sap, sbp, out=None):
    v = pred_uncond + guidance * (pred_cond - pred_uncond)
    t = lambda x: torch.tensor(x, dtype=torch.float32)   # noqa: E731  (0-dim CPU scalars: no dtype promotion, like the reference)
    x0 = t(sa) * latent - t(sb) * v
    eps = t(sa) * v + t(sb) * latent
    y = t(sap) * x0 + t(sbp) * eps
    if out is not None:
        out.copy_(y)
        return out
    return y


REFERENCE = {"gemm": _gemm, "gemm_ln": _gemm_ln, "conv3x3": _conv3x3, "groupnorm": _groupnorm, "layernorm": _layernorm, "attention": _attention,
             "temporal_attention": _temporal_attention, "cfg_ddim_step": _cfg_ddim_step}


@contextlib.contextmanager
def reference_ops(*names):
    """Swap the named ops of :mod:`imagine360_b200.ops` for torch math inside the block."""
    saved = {}
    try:
        for n in names + (("gemm_ln",) if "gemm" in names else ()):       # the folded GEMM belongs to the gemm family
            saved[n] = getattr(ops, n)
            setattr(ops, n, REFERENCE[n])
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)


def install_from_env():
    names = [n for n in os.environ.get("I360_REFERENCE_OPS", "").split(",") if n]
    for n in names + (["gemm_ln"] if "gemm" in names else []):
        setattr(ops, n, REFERENCE[n])
    return names
