"""Native forward passes of the UNet3D blocks: every tensor op is a call into ``libimagine360_b200.so``.

Activations are channels-last bf16 ``[images, H, W, C]`` with images ordered (batch, frame); the same memory is the
token matrix ``[(b f h w), C]`` of the transformer blocks, so no permute/reshape copy exists between convolutions,
spatial attention and temporal attention (the reference does one per block boundary: attention.py:266,:297;
motion_module.py:166-181,:348,:427).
"""
from __future__ import annotations

import functools
import os

import torch
import torch.nn.functional as F

from .. import ops
from .unet3d import BF16, cached, conv_w, fused_w, geglu_w, lin_w, pad8, upsample_conv_w


# ------------------------------------------------------------------------------------------------------
# layout glue for the 4/9-channel latents at the model boundary (tiny tensors; torch is plumbing here)
# ------------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------------
# NVTX ranges per block (the reference has no tracing at all).  I360_NVTX=1 names every ResnetBlock3D /
# Transformer3DModel / temporal module / WarpAttn / sampler of a step in nsys / ncu timelines; off = zero cost.
# ------------------------------------------------------------------------------------------------------
_NVTX = os.environ.get("I360_NVTX", "0") not in ("", "0")


def traced(name):
    def deco(fn):
        if not _NVTX:
            return fn

        @functools.wraps(fn)
        def wrapped(*a, **k):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco


def latents_to_nhwc(x, circular_pad: int = 0):
    """[b, c, f, h, w] -> [(b f), h, w(+2p), pad8(c)] bf16, optionally circularly padded on w (pad_pano)."""
    b, c, f, h, w = x.shape
    y = x.permute(0, 2, 3, 4, 1).reshape(b * f, h, w, c).to(BF16)
    if circular_pad:
        y = torch.cat([y[:, :, -circular_pad:], y, y[:, :, :circular_pad]], dim=2)
    if pad8(c) != c:
        y = F.pad(y, (0, pad8(c) - c))
    return y.contiguous()


def nhwc_to_latents(y, b, c):
    """[(b f), h, w, >=c] -> [b, c, f, h, w]"""
    n, h, w, _ = y.shape
    return y[..., :c].reshape(b, n // b, h, w, c).permute(0, 4, 1, 2, 3).contiguous()


def tokens(x):
    return x.view(-1, x.shape[-1])


# ------------------------------------------------------------------------------------------------------
# LayerNorm -> Linear as ONE kernel (ops.gemm_ln): the fold (W * gamma, row sums, W beta + b) is a cached weight packing, and
# the row statistics of the token matrix are written by the epilogue of the GEMM that PRODUCED it (proj_in, to_out + residual,
# ff.net[2] + residual: ``token_linear`` below), so neither a LayerNorm pass nor a normalised copy of the tokens exists.
# I360_LN_FOLD=0 restores LayerNorm + GEMM (A/B runs).  History: a first version took the statistics from the A tiles inside the
# consuming GEMM (four extra warps) and was 9 ms per step SLOWER than the separate pass (profiles/r02_bench_c3_ln_fold_*.json).
# ------------------------------------------------------------------------------------------------------
LN_FOLD = os.environ.get("I360_LN_FOLD", "1") not in ("", "0")


def token_linear(a, w, bias=None, resid=None, want_stats=True):
    """A projection whose output is the residual stream a LayerNorm reads next: -> (tokens, row statistics | None)."""
    if LN_FOLD and want_stats:
        return ops.gemm(a, w, bias=bias, resid=resid, rowstats=True)
    return ops.gemm(a, w, bias=bias, resid=resid), None


def fold_pays(k: int, act: int) -> bool:
    """Measured per consumer inside the 16x512x1024 step (profiles/r02b_ln_fold_per_shape_ab.txt, bench breakdowns): the
    folded GEGLU projection costs +0.18 ms at K = 320 and +0.07 ms at K = 640 -- its epilogue is the bottleneck there and the
    fold adds to it -- against LayerNorm passes of 0.14 / 0.07 ms, so those two keep the separate LayerNorm; everything else
    (QKV, to_q, GEGLU at K = 1280) folds."""
    return not (act == ops.ACT_GEGLU and k < 1280)


def ln_linear(t, stats, norm, owner, key, mods, bias_mod=None, act=ops.ACT_NONE, pe=None, pe_div=1, pe_mod=0):
    """act(LayerNorm(t) @ cat(W of mods)^T + bias (+ pe @ W^T per row group)).  ``pe`` fp32 [pe_mod, C] is the table
    the temporal module adds AFTER the norm (motion_module.py:350); ``stats`` are the row statistics of ``t`` from
    :func:`token_linear` (None: separate LayerNorm pass)."""
    ws = [m.weight for m in mods]
    n_total, k = sum(w.shape[0] for w in ws), ws[0].shape[1]
    geglu = act == ops.ACT_GEGLU
    if stats is not None and fold_pays(k, act) and ops.gemm_ln_supported(n_total, k, act):
        params = ws + [norm.weight, norm.bias] + ([bias_mod.bias] if bias_mod is not None and bias_mod.bias is not None else [])

        def build():
            w = torch.cat([x.to(BF16) for x in ws], 0)
            b = bias_mod.bias.to(BF16) if bias_mod is not None and bias_mod.bias is not None else None
            return ops.fold_layernorm(w, b, norm.weight.to(BF16), norm.bias.to(BF16), geglu)

        wf, u, c = cached(owner, "lnfold_" + key, params, build)
        rv = None
        if pe is not None:
            rv = cached(owner, f"lnpe_{key}_{pe.shape[0]}", ws + [pe],
                        lambda: (pe.float() @ torch.cat([x.to(BF16) for x in ws], 0).float().t()).contiguous())
        return ops.gemm_ln(t, stats, wf, u, c, norm.eps, rowvec=rv, rowvec_div=pe_div, rowvec_mod=pe_mod if pe is not None else 0, act=act)
    nrm = ops.layernorm(t, norm.weight, norm.bias, norm.eps, post_add=pe, post_div=pe_div, post_mod=pe_mod if pe is not None else 1)
    if geglu:
        wg, bg = geglu_w(owner)
        return ops.gemm(nrm, wg, bias=bg, act=ops.ACT_GEGLU)
    w = fused_w(owner, key, mods) if len(mods) > 1 else lin_w(mods[0])[0]
    b = lin_w(bias_mod)[1] if bias_mod is not None else None
    return ops.gemm(nrm, w, bias=b)


# ------------------------------------------------------------------------------------------------------
# time embedding (unet.py:718-744; MVGenModel.py:104-133)
# ------------------------------------------------------------------------------------------------------
def _mlp(t_emb, te):
    w1, b1 = lin_w(te.linear_1)
    w2, b2 = lin_w(te.linear_2)
    h = ops.gemm(t_emb, w1, bias=b1, act=ops.ACT_SILU)
    return ops.gemm(h, w2, bias=b2)


def time_embedding(unet, timesteps, fps=None):
    """-> silu(emb) [B, 4*C0] bf16, the only form the ResnetBlocks consume (resnet.py:231)."""
    emb = _mlp(unet.time_proj(timesteps).to(BF16), unet.time_embedding)
    if fps is not None:
        emb = emb + _mlp(unet.time_proj(fps).to(BF16), unet.fps_embedding)
    return F.silu(emb)


def all_temb_projections(unet, silu_emb):
    """time_emb_proj of every ResnetBlock3D of a UNet in ONE GEMM (they share the input); returns
    {resnet module: fp32 [B, Cout]} ready to be added per image inside the conv epilogue."""
    resnets = [m for m in unet.modules() if m.__class__.__name__ == "ResnetBlock3D"]

    def build():
        w = torch.cat([r.time_emb_proj.weight.to(BF16) for r in resnets], 0).contiguous()
        b = torch.cat([r.time_emb_proj.bias.to(BF16) for r in resnets], 0).contiguous()
        return w, b

    params = [p for r in resnets for p in (r.time_emb_proj.weight, r.time_emb_proj.bias)]
    w, b = cached(unet, "temb_all", params, build)
    out = ops.gemm(silu_emb, w, bias=b).float()
    res, o = {}, 0
    for r in resnets:
        n = r.out_channels
        res[r] = out[:, o:o + n].contiguous()
        o += n
    return res


# ------------------------------------------------------------------------------------------------------
# ResnetBlock3D (resnet.py:221-254), optionally on the pano circular halo (MVGenModel.py:276-281)
# ------------------------------------------------------------------------------------------------------
# conv -> GroupNorm (opt-in, I360_CONV_GN_STATS=1): the conv epilogue accumulates per-channel statistics of what it stores and the
# GroupNorm that reads the tensor folds channels into groups instead of running a statistics pass (norm2 after conv1;
# Transformer3DModel.norm after a block's conv2).  Measured on the C3 step (profiles/r02b_conv_gn_chanstats_ab.txt): the 68
# statistics passes it removes are 3.3 ms, the conv epilogues pay 1.7 ms and the apply kernels 1.0 ms for it -> -0.4 ms of
# ~300 ms, inside the run-to-run noise, while the engine's own time (the roofline kernel) grows.  Off by default.
CONV_GN = os.environ.get("I360_CONV_GN_STATS", "0") not in ("", "0")


def _chan_stats_ok(h: int, w: int) -> bool:
    return CONV_GN and h * w >= 32          # a warp's 32 tile rows must lie in one image


@traced("resnet_block")
def resnet_block(x, r, temb, frames: int, groups: int, skip=None, halo: int = 0, out_stats: bool = False):
    """x [N,H,W,C1] (+ skip [N,H,W,C2] = the torch.cat of the up path) -> [N,H,W,Cout].
    halo > 0: pad_pano(halo) -> block -> unpad_pano(halo): the GroupNorm statistics are taken over the padded tensor
    (norm2 even over conv1's zero-padded outer columns) exactly as the reference does.
    ``out_stats``: also leave the per-channel statistics of the result on it (``_i360_chan_stats``) for the GroupNorm of the
    Transformer3DModel that follows."""
    eps = r.norm1.eps
    h = ops.groupnorm(x, r.norm1.weight, r.norm1.bias, groups, eps, True, x2=skip, pad=halo)
    w1, b1 = conv_w(r.conv1)
    if _chan_stats_ok(h.shape[1], h.shape[2]):
        h, st = ops.conv3x3(h, w1, bias=b1, rowvec=temb, rowvec_div=frames, chan_stats=True)
        h = ops.groupnorm(h, r.norm2.weight, r.norm2.bias, groups, eps, True, chan_stats=st)
    else:
        h = ops.conv3x3(h, w1, bias=b1, rowvec=temb, rowvec_div=frames)
        h = ops.groupnorm(h, r.norm2.weight, r.norm2.bias, groups, eps, True)
    want = out_stats and _chan_stats_ok(x.shape[1], x.shape[2])
    if r.conv_shortcut is not None:
        w2, b2 = conv_w(r.conv2, shortcut=r.conv_shortcut)
        y = ops.conv3x3(h, w2, bias=b2, x2=x, x3=skip, crop=halo, out_scale=1.0 / r.output_scale_factor, chan_stats=want)
    else:
        assert skip is None
        w2, b2 = conv_w(r.conv2)
        y = ops.conv3x3(h, w2, bias=b2, resid=x, crop=halo, out_scale=1.0 / r.output_scale_factor, chan_stats=want)
    if want:
        y, st = y
        y._i360_chan_stats = st
    return y


# I360_CONV_S2_IM2COL=1: the round-1 path (materialised im2col + GEMM) for A/B runs
S2_IM2COL = os.environ.get("I360_CONV_S2_IM2COL", "0") not in ("", "0")


@traced("downsample")
def downsample(x, d, circular: bool):
    """Downsample3D (resnet.py:132-140); circular: pad_pano(2) -> conv -> unpad_pano(1) (MVGenModel.py:305-314).
    Implicit GEMM over TMA boxes with traversal stride 2: no im2col buffer; the circular case materialises the two halo columns
    per side of the input (one small copy) and crops one output column per side in the kernel."""
    n, h, w, c = x.shape
    wp, b = conv_w(d.conv)
    if S2_IM2COL:
        y = ops.gemm(ops.im2col_s2(x, circular), wp, bias=b)
        return y.view(n, h // 2, w // 2, -1)
    if circular:
        xp = torch.cat([x[:, :, -2:], x, x[:, :, :2]], dim=2)
        return ops.conv3x3_s2(xp, wp, b, pad_lo=1, crop=1)
    return ops.conv3x3_s2(x, wp, b, pad_lo=1)


# I360_UPSAMPLE_SUBPIXEL=0: materialise the upsampled tensor and run the plain 3x3 conv on it (A/B runs)
SUBPIXEL = os.environ.get("I360_UPSAMPLE_SUBPIXEL", "1") not in ("", "0")


@traced("upsample")
def upsample(x, u, circular: bool):
    """Upsample3D (resnet.py:86-114); circular: pad_pano(1) -> x2 -> conv -> unpad_pano(2) (MVGenModel.py:449-456).
    Sub-pixel form: four 2x2-tap convolutions of the low-resolution tensor (one per output parity) with pre-summed weights --
    no upsampled tensor, 4/9 of the multiply-adds; the circular halo is one low-resolution column per side, cropped by the
    kernel (the only copy is that small padded low-resolution tensor)."""
    if SUBPIXEL:
        weff, b = upsample_conv_w(u.conv)
        if circular:
            xp = torch.cat([x[:, :, -1:], x, x[:, :, :1]], dim=2)
            return ops.conv_upsample2x(xp, weff, b, crop=1)
        return ops.conv_upsample2x(x, weff, b)
    wp, b = conv_w(u.conv)
    if circular:
        return ops.conv3x3(ops.upsample2x(x, pad_in=1), wp, bias=b, crop=2)
    return ops.conv3x3(ops.upsample2x(x), wp, bias=b)


# ------------------------------------------------------------------------------------------------------
# Transformer3DModel (attention.py:246-301) with BasicTransformerBlock (:461-508)
# ------------------------------------------------------------------------------------------------------
class Context:
    """Cross-attention context of one branch for one step: text tokens (step-invariant) and image tokens
    (re-noised every step).  K/V projections are computed once per batch element instead of once per frame
    (the reference repeats the context ``f`` times, attention.py:257)."""

    def __init__(self, text, ip):
        self.text, self.ip = text.contiguous(), ip.contiguous()   # [Bc, 77, D], [Bc, n_ip, D]
        self.n_ctx = text.shape[0]


@traced("spatial_transformer")
def spatial_transformer(x, t3d, ctx: Context, frames: int):
    n, h, w, c = x.shape
    heads, hd = t3d.heads, t3d.dim_head
    npix = h * w
    # statistics of x from the epilogue of the conv that produced it (resnet_block(..., out_stats=True)), when it left them
    cst = getattr(x, "_i360_chan_stats", None)
    hn = ops.groupnorm(x, t3d.norm.weight, t3d.norm.bias, t3d.groups, 1e-6, False, chan_stats=cst)
    wi, bi = lin_w(t3d.proj_in)
    t, st = token_linear(tokens(hn), wi, bi)
    for blk in t3d.transformer_blocks:
        # --- attn1: fused QKV projection, flash attention reading q/k/v as column slices ---
        a1 = blk.attn1
        qkv = ln_linear(t, st, blk.norm1, a1, "qkv", [a1.to_q, a1.to_k, a1.to_v])
        o = torch.empty_like(t)
        ops.attention(ops.seq_view(qkv, n, npix, 0), ops.seq_view(qkv, n, npix, c), ops.seq_view(qkv, n, npix, 2 * c),
                      ops.seq_view(o, n, npix), heads, hd, n)
        wo, bo = lin_w(a1.to_out[0])
        t, st = token_linear(o, wo, bo, t)
        # --- attn2: text + image-prompt cross attention, outputs summed before to_out (attention.py:148) ---
        a2 = blk.attn2
        q = ln_linear(t, st, blk.norm2, a2, "q", [a2.to_q])
        nt, ni = ctx.text.shape[1], ctx.ip.shape[1]
        kv_t = ops.gemm(ctx.text.view(-1, ctx.text.shape[-1]), fused_w(a2, "kv", [a2.to_k, a2.to_v]))
        ip = ctx.ip[..., : a2.image_cross_attention_dim] if a2.image_cross_attention_dim != a2.cross_attention_dim else ctx.ip
        kv_i = ops.gemm(ip.reshape(-1, ip.shape[-1]), fused_w(a2, "kv_ip", [a2.to_k_ip, a2.to_v_ip]))
        o = torch.empty_like(t)
        if a2.ip_scale != 1.0:
            raise NotImplementedError("ip scale != 1.0")
        if ops.cross_attention_text_ip_supported(hd, nt, ni) and n == ctx.n_ctx * frames:
            # both key sets stay in shared memory; one read of q, one write of o
            ops.cross_attention_text_ip(q, o, kv_t, nt, kv_i, ni, ctx.n_ctx, heads, hd)
        else:
            qv, ov = ops.seq_view(q, n, npix), ops.seq_view(o, n, npix)
            ops.attention(qv, ops.seq_view(kv_t, ctx.n_ctx, nt, 0, share_div=frames),
                          ops.seq_view(kv_t, ctx.n_ctx, nt, c, share_div=frames), ov, heads, hd, n)
            ops.attention(qv, ops.seq_view(kv_i, ctx.n_ctx, ni, 0, share_div=frames),
                          ops.seq_view(kv_i, ctx.n_ctx, ni, c, share_div=frames), ov, heads, hd, n, accumulate=True)
        wo, bo = lin_w(a2.to_out[0])
        t, st = token_linear(o, wo, bo, t, want_stats=fold_pays(c, ops.ACT_GEGLU))
        # --- GEGLU feed-forward ---
        t, st = feed_forward(t, st, blk.ff, blk.norm3)
    wo, bo = lin_w(t3d.proj_out)
    return ops.gemm(t, wo, bias=bo, resid=tokens(x)).view(n, h, w, c)


@traced("feed_forward")
def feed_forward(t, st, ff, norm):
    """-> (tokens, row statistics of the result): the next block's first LayerNorm reads them."""
    g = ln_linear(t, st, norm, ff, "geglu", [ff.net[0].proj], bias_mod=ff.net[0].proj, act=ops.ACT_GEGLU)
    w2, b2 = lin_w(ff.net[2])
    return token_linear(g, w2, b2, t)


# ------------------------------------------------------------------------------------------------------
# VanillaTemporalModule (motion_module.py:158-185, :247-259, :343-429)
# ------------------------------------------------------------------------------------------------------
def _pe_table(att, frames, dtype_like):
    """pos_encoder.pe[:F] as fp32 values of the (bf16-cast) buffer the reference adds (motion_module.py:350)."""
    pe = att.pos_encoder.pe
    return cached(att, f"pe{frames}", [pe], lambda: pe[0, :frames].to(dtype_like).float().contiguous())


@traced("temporal_module")
def temporal_module(x, mm, frames: int):
    tt = mm.temporal_transformer
    n, h, w, c = x.shape
    d = h * w
    hn = ops.groupnorm(x, tt.norm.weight, tt.norm.bias, tt.groups, 1e-6, False)
    wi, bi = lin_w(tt.proj_in)
    t, st = token_linear(tokens(hn), wi, bi)
    for blk in tt.transformer_blocks:
        n_att = len(blk.attention_blocks)
        for ai, (att, norm) in enumerate(zip(blk.attention_blocks, blk.norms)):
            qkv = ln_linear(t, st, norm, att, "qkv", [att.to_q, att.to_k, att.to_v], pe=_pe_table(att, frames, BF16), pe_div=d,
                            pe_mod=frames)
            o = torch.empty_like(t)
            ops.temporal_attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], o, n // frames, frames, d, tt.heads, tt.dim_head)
            wo, bo = lin_w(att.to_out[0])
            # the last attention block's output is read by ff_norm -> GEGLU, the others by the next block's norm -> QKV
            t, st = token_linear(o, wo, bo, t, want_stats=ai + 1 < n_att or fold_pays(c, ops.ACT_GEGLU))
        t, st = feed_forward(t, st, blk.ff, blk.ff_norm)
    wo, bo = lin_w(tt.proj_out)
    return ops.gemm(t, wo, bias=bo, resid=tokens(x)).view(n, h, w, c)


# ------------------------------------------------------------------------------------------------------
# conv_in / conv_out
# ------------------------------------------------------------------------------------------------------
@traced("conv_in")
def conv_in(unet, latents, circular: bool):
    """[b, 9, f, h, w] -> [(b f), h, w, C0]; pano: pad_pano(1) -> conv -> unpad_pano(1) (MVGenModel.py:136-143)."""
    x = latents_to_nhwc(latents, 1 if circular else 0)
    wp, b = conv_w(unet.conv_in, cin_pad=x.shape[-1])
    return ops.conv3x3(x, wp, bias=b, crop=1 if circular else 0)


@traced("conv_out")
def conv_out(unet, x, batch: int, circular: bool):
    """GroupNorm -> SiLU -> conv_out; pano: the norm runs BEFORE pad_pano(1) (MVGenModel.py:472-478)."""
    co = unet.conv_out.out_channels
    h = ops.groupnorm(x, unet.conv_norm_out.weight, unet.conv_norm_out.bias, unet.groups, unet.conv_norm_out.eps, True,
                      pad=1 if circular else 0, stats_pad=0)
    wp, b = conv_w(unet.conv_out, cout_pad=pad8(co))
    y = ops.conv3x3(h, wp, bias=b, crop=1 if circular else 0)
    return nhwc_to_latents(y, batch, co)


# ------------------------------------------------------------------------------------------------------
# single-branch forward (unet.py:632-856) -- configs C1 / C2
# ------------------------------------------------------------------------------------------------------
def unet_single_forward(unet, sample, timestep, ctx_tokens, fps=None):
    b, _, frames, _, _ = sample.shape
    dev = sample.device
    t = timestep.reshape(-1).to(dev).expand(b)
    silu_emb = time_embedding(unet, t, None if fps is None else fps.reshape(-1).to(dev).expand(b))
    temb = all_temb_projections(unet, silu_emb)
    nt = ctx_tokens.shape[1] - unet.num_tokens
    ctx = Context(ctx_tokens[:, :nt].to(BF16), ctx_tokens[:, nt:].to(BF16))
    g = unet.groups
    x = conv_in(unet, sample, False)
    skips = [x]
    for blk in unet.down_blocks:
        for j, r in enumerate(blk.resnets):
            x = resnet_block(x, r, temb[r], frames, g, out_stats=blk.has_cross_attention)
            if blk.has_cross_attention:
                x = spatial_transformer(x, blk.attentions[j], ctx, frames)
            if blk.motion_modules[j] is not None:      # UNet3DConditionModel.forward runs them in every block
                x = temporal_module(x, blk.motion_modules[j], frames)
            skips.append(x)
        if blk.downsamplers is not None:
            x = downsample(x, blk.downsamplers[0], False)
            skips.append(x)
    mid = unet.mid_block
    x = resnet_block(x, mid.resnets[0], temb[mid.resnets[0]], frames, g, out_stats=True)
    for i, att in enumerate(mid.attentions):
        x = spatial_transformer(x, att, ctx, frames)
        if mid.motion_modules[i] is not None:
            x = temporal_module(x, mid.motion_modules[i], frames)
        x = resnet_block(x, mid.resnets[i + 1], temb[mid.resnets[i + 1]], frames, g)
    for blk in unet.up_blocks:
        for j, r in enumerate(blk.resnets):
            x = resnet_block(x, r, temb[r], frames, g, skip=skips.pop(), out_stats=blk.has_cross_attention)
            if blk.has_cross_attention:
                x = spatial_transformer(x, blk.attentions[j], ctx, frames)
            if blk.motion_modules[j] is not None:
                x = temporal_module(x, blk.motion_modules[j], frames)
        if blk.upsamplers is not None:
            x = upsample(x, blk.upsamplers[0], False)
    return conv_out(unet, x, b, False)
