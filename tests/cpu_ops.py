"""Test infrastructure: torch emulations of the documented contract of EVERY ``imagine360_b200.ops`` entry point the UNet /
MultiViewBaseModel host code calls, so that the host wiring (weight packing, token views, op order, halo / crop bookkeeping,
caches) can be executed and checked against the oracle on a machine without a GPU.  The kernels themselves are checked in
the ``-m gpu`` tests; nothing in the product imports this file (the product fails loudly without the CUDA library).

The GEMM / conv / norm / attention emulations are the bisecting aids of ``imagine360_b200/debug.py``; the rest is here."""
import contextlib

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def _cross_attention_text_ip(q, out, kv_text, nt, kv_ip, ni, n_ctx, heads, head_dim, scale=None):
    c = heads * head_dim
    rows = q.shape[0] // n_ctx
    qh = q[:, :c].float().view(n_ctx, rows, heads, head_dim).transpose(1, 2)

    def branch(kv, n):
        k = kv[:, :c].float().view(n_ctx, n, heads, head_dim).transpose(1, 2)
        v = kv[:, c:2 * c].float().view(n_ctx, n, heads, head_dim).transpose(1, 2)
        return F.scaled_dot_product_attention(qh, k, v, scale=scale)

    y = branch(kv_text, nt) + branch(kv_ip, ni)                # one fp32 sum, one rounding
    out[:, :c] = y.transpose(1, 2).reshape(n_ctx * rows, c).to(BF16)


def _conv_upsample2x(x, w_eff, bias=None, crop=0, gn_groups=None):
    B, H, W, Cin = x.shape
    Cout = w_eff.shape[1]
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))
    out = torch.zeros(B, Cout, 2 * H, 2 * W)
    for a in (0, 1):
        for b in (0, 1):
            w4 = w_eff[a * 2 + b].float().view(Cout, 4, Cin)
            acc = 0
            for dr in (0, 1):
                for dc in (0, 1):
                    acc = acc + torch.einsum("bchw,oc->bohw", xp[:, :, a + dr:a + dr + H, b + dc:b + dc + W], w4[:, dr * 2 + dc])
            out[:, :, a::2, b::2] = acc
    if bias is not None:
        out = out + bias.float()[None, :, None, None]
    if crop:
        out = out[..., 2 * crop:-2 * crop]
    y = out.permute(0, 2, 3, 1).to(BF16).contiguous()
    if gn_groups is not None:
        v = y.double().view(B, -1, gn_groups, Cout // gn_groups)
        return y, torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], -1)
    return y


def _conv3x3_s2(x, w_packed, bias=None, pad_lo=1, crop=0):
    B, H, W, Cin = x.shape
    w = w_packed.float().view(-1, 3, 3, Cin).permute(0, 3, 1, 2)
    xin = x.float().permute(0, 3, 1, 2)
    if pad_lo == 0:
        xin = F.pad(xin, (0, 1, 0, 1))
    y = F.conv2d(xin, w, None if bias is None else bias.float(), stride=2, padding=pad_lo)
    if crop:
        y = y[..., crop:-crop]
    return y.permute(0, 2, 3, 1).to(BF16).contiguous()


def _upsample2x(x, pad_in=0):
    if pad_in:
        x = torch.cat([x[:, :, -pad_in:], x, x[:, :, :pad_in]], 2)
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous()


def _grid_sample(img, grid, nearest=False):
    return F.grid_sample(img, grid, mode="nearest" if nearest else "bilinear", padding_mode="zeros", align_corners=True)


def _avgpool_frames4(x):
    B, Fr, D, C = x.shape
    return x.float().view(B, Fr // 4, 4, D, C).mean(2).to(BF16)


def _axpby(x, y, a, b):
    # reference order (add_noise_to_condition, MVGenModel.py:11-14): (y * b) rounded to bf16, then a * x + that
    t = (y.float() * b).to(BF16).float() if y is not None else 0.0
    return (a * x.float() + t).to(BF16)


def _im2col_s2(x, circular, pad_lo=1):
    B, H, W, C = x.shape
    Ho, Wo = H // 2, W // 2
    out = torch.zeros(B, Ho, Wo, 9, C, dtype=x.dtype)
    ho, wo = torch.arange(Ho), torch.arange(Wo)
    for tap in range(9):
        hi, wi = 2 * ho + tap // 3 - pad_lo, 2 * wo + tap % 3 - pad_lo
        okh = (hi >= 0) & (hi < H)
        if circular:
            wi, okw = (wi + W) % W, torch.ones(Wo, dtype=torch.bool)
        else:
            okw = (wi >= 0) & (wi < W)
        v = x[:, hi.clamp(0, H - 1)][:, :, wi.clamp(0, W - 1)]
        out[:, :, :, tap] = v * (okh[:, None] & okw[None, :])[None, :, :, None].to(x.dtype)
    return out.view(B * Ho * Wo, 9 * C)


def _softmax_rows(x):
    return x.float().softmax(-1).to(BF16)


@contextlib.contextmanager
def cpu_ops(calls=None):
    """Every kernel entry point of ``imagine360_b200.ops`` used by the UNet / MVGen host code -> its torch emulation.
    ``calls`` (a dict) receives the number of calls per entry point (``conv3x3+chan_stats`` / ``groupnorm+chan_stats`` /
    ``groupnorm+stats`` count the calls on the conv -> GroupNorm routes), so that a test can tell which sequence of C-ABI calls the host issued."""
    import collections
    import functools
    from imagine360_b200 import debug, ops
    table = dict(debug.REFERENCE)
    table.update(cross_attention_text_ip=_cross_attention_text_ip, conv_upsample2x=_conv_upsample2x, conv3x3_s2=_conv3x3_s2,
                 upsample2x=_upsample2x, im2col_s2=_im2col_s2, grid_sample=_grid_sample, avgpool_frames4=_avgpool_frames4, axpby=_axpby,
                 softmax_rows=_softmax_rows)
    counts = collections.Counter()

    def counted(name, f):
        @functools.wraps(f)
        def g(*a, **kw):
            counts[name] += 1
            if kw.get("chan_stats") is not None and kw.get("chan_stats") is not False:
                counts[name + "+chan_stats"] += 1
            if kw.get("stats") is not None:
                counts[name + "+stats"] += 1
            return f(*a, **kw)
        return g

    table = {k: counted(k, f) for k, f in table.items()}
    table["on_device"] = lambda device: contextlib.nullcontext()
    saved = {k: getattr(ops, k) for k in table}
    for k, f in table.items():
        setattr(ops, k, f)
    try:
        yield
    finally:
        for k, f in saved.items():
            setattr(ops, k, f)
        if calls is not None:
            calls.update(counts)
