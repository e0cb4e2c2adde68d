"""Oracle: UNet3DConditionModel building blocks and single-branch forward.

Follows animatediff/models/{resnet,attention,motion_module,unet_blocks,unet}.py of the reference.
Tensors are the reference's layouts: feature maps [b, c, f, h, w], tokens [(b f), n, c].
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .nn_ops import (P, attention_core, conv2d_frames, geglu_ff, group_norm_frames, linear, sinusoid_table,
                     timestep_embedding, timestep_mlp)

DEFAULT_CFG = dict(
    groups=32,                 # norm_num_groups
    resnet_eps=1e-5,           # norm_eps (unet.py:90)
    heads=(5, 10, 20, 20),     # attention_head_dim == head COUNT in this API (unet_blocks.py:413-416)
    mm_heads=8,                # motion_module_kwargs.num_attention_heads (configs/prompt-dual.yaml:23)
    num_tokens=64,             # IP tokens (yaml:41)
    ip_scale=1.0,
    temporal_pe_max_len=64,
    time_dim=None,             # block_out_channels[0]; inferred from conv_in
    adapter_heads=12, adapter_dim_head=64, tproj_heads=8, tproj_dim_head=64,
)


# --------------------------------------------------------------------------------------------
# ResnetBlock3D (animatediff/models/resnet.py:221-254)
# --------------------------------------------------------------------------------------------
def resnet_block(x, temb, p: P, cfg):
    g, eps = cfg["groups"], cfg["resnet_eps"]
    h = F.silu(group_norm_frames(x, p, "norm1", g, eps))
    h = conv2d_frames(h, p, "conv1")
    if temb is not None:
        t = linear(F.silu(temb), p, "time_emb_proj")[:, :, None, None, None]
        h = h + t
    h = F.silu(group_norm_frames(h, p, "norm2", g, eps))
    h = conv2d_frames(h, p, "conv2")
    if p.has("conv_shortcut.weight"):
        x = conv2d_frames(x, p, "conv_shortcut", padding=0)
    return x + h  # output_scale_factor == 1


def downsample(x, p: P):
    """Downsample3D: stride-2 3x3 conv, padding 1 (resnet.py:117-140)."""
    return conv2d_frames(x, p, "conv", stride=2, padding=1)


def upsample(x, p: P):
    """Upsample3D: nearest x2 on (h, w) then 3x3 conv (resnet.py:86-114)."""
    b, c, f, h, w = x.shape
    y = F.interpolate(x.float(), scale_factor=[1.0, 2.0, 2.0], mode="nearest").to(x.dtype)
    return conv2d_frames(y, p, "conv")


# --------------------------------------------------------------------------------------------
# Transformer3DModel / BasicTransformerBlock / IPCrossAttention (animatediff/models/attention.py)
# --------------------------------------------------------------------------------------------
def self_attention(x, p: P, heads):
    """attn1: Attention + XFormersAttnProcessor (attention_processor.py:1210-1283)."""
    q, k, v = linear(x, p, "to_q"), linear(x, p, "to_k"), linear(x, p, "to_v")
    return linear(attention_core(q, k, v, heads), p, "to_out.0")


def ip_cross_attention(x, ctx, p: P, heads, num_tokens, ip_scale):
    """IPCrossAttention.forward (attention.py:65-156): the last ``num_tokens`` context rows are
    the image tokens; two attentions are summed BEFORE to_out."""
    end = ctx.shape[1] - num_tokens
    text, ip = ctx[:, :end], ctx[:, end:]
    q = linear(x, p, "to_q")
    o = attention_core(q, linear(text, p, "to_k"), linear(text, p, "to_v"), heads)
    o_ip = attention_core(q, linear(ip, p, "to_k_ip"), linear(ip, p, "to_v_ip"), heads)
    return linear(o + ip_scale * o_ip, p, "to_out.0")


def spatial_transformer(x, ctx, p: P, heads, cfg):
    """Transformer3DModel.forward with use_linear_projection=True (attention.py:246-301)."""
    b, c, f, h, w = x.shape
    frames = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    ctx_f = ctx.repeat_interleave(f, dim=0)  # 'b n c -> (b f) n c'
    y = F.group_norm(frames, cfg["groups"], p["norm.weight"], p["norm.bias"], 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    y = linear(y, p, "proj_in")
    for i in range(p.count("transformer_blocks")):
        bp = p.sub(f"transformer_blocks.{i}")
        y = self_attention(F.layer_norm(y, (c,), bp["norm1.weight"], bp["norm1.bias"]), bp.sub("attn1"), heads) + y
        y = ip_cross_attention(F.layer_norm(y, (c,), bp["norm2.weight"], bp["norm2.bias"]), ctx_f, bp.sub("attn2"),
                               heads, cfg["num_tokens"], cfg["ip_scale"]) + y
        y = geglu_ff(F.layer_norm(y, (c,), bp["norm3.weight"], bp["norm3.bias"]), bp.sub("ff")) + y
    y = linear(y, p, "proj_out")
    y = y.reshape(b * f, h, w, c).permute(0, 3, 1, 2) + frames
    return y.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


# --------------------------------------------------------------------------------------------
# VanillaTemporalModule (animatediff/models/motion_module.py:52-429)
# --------------------------------------------------------------------------------------------
def temporal_self_attention(x, f, p: P, heads, cfg, math_path=True):
    """VersatileAttention.forward (motion_module.py:343-429).  x: [(b f), d, c] -> same.
    Sinusoidal PE is added to the hidden states that feed q, k AND v (:350)."""
    bf, d, c = x.shape
    b = bf // f
    y = x.reshape(b, f, d, c).permute(0, 2, 1, 3).reshape(b * d, f, c)  # '(b f) d c -> (b d) f c'
    pe = p.get("pos_encoder.pe")
    pe = pe[0, :f] if pe is not None else sinusoid_table(cfg["temporal_pe_max_len"], c, x.device)[:f]
    y = y + pe.to(y.dtype)
    q, k, v = linear(y, p, "to_q"), linear(y, p, "to_k"), linear(y, p, "to_v")
    o = attention_core(q, k, v, heads, math_path=math_path)
    o = linear(o, p, "to_out.0")
    return o.reshape(b, d, f, c).permute(0, 2, 1, 3).reshape(bf, d, c)


def temporal_module(x, p: P, cfg):
    """TemporalTransformer3DModel.forward (motion_module.py:158-185)."""
    p = p.sub("temporal_transformer")
    b, c, f, h, w = x.shape
    frames = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    y = F.group_norm(frames, cfg["groups"], p["norm.weight"], p["norm.bias"], 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    y = linear(y, p, "proj_in")
    for i in range(p.count("transformer_blocks")):
        bp = p.sub(f"transformer_blocks.{i}")
        for j in range(bp.count("attention_blocks")):
            n = F.layer_norm(y, (c,), bp[f"norms.{j}.weight"], bp[f"norms.{j}.bias"])
            y = temporal_self_attention(n, f, bp.sub(f"attention_blocks.{j}"), cfg["mm_heads"], cfg) + y
        y = geglu_ff(F.layer_norm(y, (c,), bp["ff_norm.weight"], bp["ff_norm.bias"]), bp.sub("ff")) + y
    y = linear(y, p, "proj_out")
    y = y.reshape(b * f, h, w, c).permute(0, 3, 1, 2) + frames
    return y.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


# --------------------------------------------------------------------------------------------
# adapter: TemporalProjection + Resampler (animatediff/models/resampler.py)
# --------------------------------------------------------------------------------------------
def _legacy_self_attention(x, p: P, heads):
    """diffusers.models.attention.CrossAttention, math path (diffusers/models/attention.py:510-673)."""
    q, k, v = linear(x, p, "to_q"), linear(x, p, "to_k"), linear(x, p, "to_v")
    return linear(attention_core(q, k, v, heads, math_path=True), p, "to_out.0")


def _plain_ff(x, p: P):
    """resampler.py:15-22: Sequential(LayerNorm, Linear(no bias), GELU, Linear(no bias))."""
    c = x.shape[-1]
    y = F.layer_norm(x, (c,), p["0.weight"], p["0.bias"])
    return F.linear(F.gelu(F.linear(y, p["1.weight"])), p["3.weight"])


def temporal_projection(feats, p: P, cfg):
    """TemporalProjection.forward (resampler.py:231-267).  feats: [b, f, hw, c] SAM features."""
    b, f, d, c = feats.shape
    hs = int(d ** 0.5)
    x = feats.reshape(b * f, hs, hs, c).permute(0, 3, 1, 2)
    x = F.conv2d(x, p["patch_embed.weight"], p["patch_embed.bias"], stride=4)  # spacial_compress (dim < 1024)
    c = x.shape[1]
    x = x.flatten(2).permute(0, 2, 1).reshape(b, f, -1, c)
    d = x.shape[2]
    heads = cfg["tproj_heads"]

    def temporal_stage(x, attn, norm, ff, ffnorm):
        b, f, d, c = x.shape
        y = x.permute(0, 2, 1, 3).reshape(b * d, f, c)
        y = _legacy_self_attention(F.layer_norm(y, (c,), p[f"{norm}.weight"], p[f"{norm}.bias"]), p.sub(attn), heads) + y
        x = y.reshape(b, d, f, c).permute(0, 2, 1, 3)
        x = _plain_ff(F.layer_norm(x, (c,), p[f"{ffnorm}.weight"], p[f"{ffnorm}.bias"]), p.sub(ff)) + x
        # avg_pool1d(kernel 4) over frames
        y = x.permute(0, 2, 3, 1).reshape(b * d, c, f)
        y = F.avg_pool1d(y, kernel_size=4)
        return y.reshape(b, d, c, -1).permute(0, 3, 1, 2)

    x = temporal_stage(x, "attn_temp", "norm_temp", "ff", "norm1")
    x = temporal_stage(x, "attn_temp_2", "norm_temp_2", "ff_2", "norm2")
    return x  # [b, f/16, d, c]


def resampler(x, p: P, cfg):
    """Resampler.forward with PerceiverAttention (resampler.py:36-160)."""
    heads, dh = cfg["adapter_heads"], cfg["adapter_dim_head"]
    lat = p["latents"].repeat(x.shape[0], 1, 1)
    x = linear(x, p, "proj_in")
    dim = x.shape[-1]
    for i in range(p.count("layers")):
        ap, fp = p.sub(f"layers.{i}.0"), p.sub(f"layers.{i}.1")
        xn = F.layer_norm(x, (dim,), ap["norm1.weight"], ap["norm1.bias"])
        ln = F.layer_norm(lat, (dim,), ap["norm2.weight"], ap["norm2.bias"])
        q = F.linear(ln, ap["to_q.weight"])
        kv = F.linear(torch.cat([xn, ln], dim=-2), ap["to_kv.weight"])
        k, v = kv.chunk(2, dim=-1)
        b, l, _ = q.shape
        qh = q.reshape(b, l, heads, dh).transpose(1, 2)
        kh = k.reshape(b, -1, heads, dh).transpose(1, 2)
        vh = v.reshape(b, -1, heads, dh).transpose(1, 2)
        s = dh ** -0.25
        wgt = torch.softmax(((qh * s) @ (kh * s).transpose(-2, -1)).float(), dim=-1).to(qh.dtype)
        o = (wgt @ vh).permute(0, 2, 1, 3).reshape(b, l, -1)
        lat = F.linear(o, ap["to_out.weight"]) + lat
        lat = _plain_ff(lat, fp) + lat
    lat = linear(lat, p, "proj_out")
    return F.layer_norm(lat, (lat.shape[-1],), p["norm_out.weight"], p["norm_out.bias"])


# --------------------------------------------------------------------------------------------
# time / fps embedding
# --------------------------------------------------------------------------------------------
def time_embedding(timesteps, fps, p: P, dtype, time_dim):
    """unet.py:718-744 / MVGenModel.py:104-133: emb = time_embedding(time_proj(t)) (+ fps_embedding(time_proj(fps)))."""
    emb = timestep_mlp(timestep_embedding(timesteps, time_dim).to(dtype), p.sub("time_embedding"))
    if fps is not None:
        emb = emb + timestep_mlp(timestep_embedding(fps, time_dim).to(dtype), p.sub("fps_embedding"))
    return emb


# --------------------------------------------------------------------------------------------
# full single-branch forward (animatediff/models/unet.py:632-856) -- config C1/C2 oracle
# --------------------------------------------------------------------------------------------
def unet3d_forward(sd, sample, timestep, ctx, cfg=None, fps=None):
    """UNet3DConditionModel.forward with use_ip_plus_cross_attention=False: ``ctx`` must already hold
    text + num_tokens image rows (SURVEY.md §8(a) note on C1).  Unlike the dual-branch forward this
    one DOES run the motion modules of DownBlock3D / UpBlock3D (unet_blocks.py:587,:843)."""
    cfg = {**DEFAULT_CFG, **(cfg or {})}
    p = P(sd)
    dtype = sample.dtype
    time_dim = p["conv_in.weight"].shape[0]
    t = timestep.reshape(-1).expand(sample.shape[0])
    emb = time_embedding(t, None if fps is None else fps.expand(sample.shape[0]), p, dtype, time_dim)
    x = conv2d_frames(sample, p, "conv_in")
    skips = [x]
    nd = p.count("down_blocks")
    for i in range(nd):
        bp = p.sub(f"down_blocks.{i}")
        has_attn = bp.has_prefix("attentions")
        for j in range(bp.count("resnets")):
            x = resnet_block(x, emb, bp.sub(f"resnets.{j}"), cfg)
            if has_attn:
                x = spatial_transformer(x, ctx, bp.sub(f"attentions.{j}"), cfg["heads"][i], cfg)
            if bp.has_prefix(f"motion_modules.{j}"):
                x = temporal_module(x, bp.sub(f"motion_modules.{j}"), cfg)
            skips.append(x)
        if bp.has_prefix("downsamplers"):
            x = downsample(x, bp.sub("downsamplers.0"))
            skips.append(x)
    mp = p.sub("mid_block")
    x = resnet_block(x, emb, mp.sub("resnets.0"), cfg)
    for i in range(mp.count("attentions")):
        x = spatial_transformer(x, ctx, mp.sub(f"attentions.{i}"), cfg["heads"][-1], cfg)
        if mp.has_prefix(f"motion_modules.{i}"):
            x = temporal_module(x, mp.sub(f"motion_modules.{i}"), cfg)
        x = resnet_block(x, emb, mp.sub(f"resnets.{i + 1}"), cfg)
    nu = p.count("up_blocks")
    for i in range(nu):
        bp = p.sub(f"up_blocks.{i}")
        has_attn = bp.has_prefix("attentions")
        for j in range(bp.count("resnets")):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(x, emb, bp.sub(f"resnets.{j}"), cfg)
            if has_attn:
                x = spatial_transformer(x, ctx, bp.sub(f"attentions.{j}"), cfg["heads"][nu - 1 - i], cfg)
            if bp.has_prefix(f"motion_modules.{j}"):
                x = temporal_module(x, bp.sub(f"motion_modules.{j}"), cfg)
        if bp.has_prefix("upsamplers"):
            x = upsample(x, bp.sub("upsamplers.0"))
    x = F.silu(group_norm_frames(x, p, "conv_norm_out", cfg["groups"], cfg["resnet_eps"]))
    return conv2d_frames(x, p, "conv_out")
