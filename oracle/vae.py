"""Oracle: AutoencoderKL encode/decode (SD-2.1 VAE topology).

Follows diffusers/models/vae.py:67-224 (Encoder/Decoder), :501-610 (AutoencoderKL),
diffusers/models/resnet.py:367-496 (ResnetBlock2D), :77-190 (Upsample2D/Downsample2D),
diffusers/models/attention.py:247-380 (single-head AttentionBlock, fp32 softmax),
diffusers/models/unet_2d_blocks.py:320-400 (UNetMidBlock2D)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .nn_ops import P, linear

GROUPS = 32
EPS = 1e-6


def _gn(x, p, name, groups=GROUPS):
    return F.group_norm(x, groups, p[f"{name}.weight"], p[f"{name}.bias"], EPS)


def _conv(x, p, name, stride=1, padding=1):
    return F.conv2d(x, p[f"{name}.weight"], p[f"{name}.bias"], stride=stride, padding=padding)


def resnet2d(x, p: P, groups=GROUPS):
    h = _conv(F.silu(_gn(x, p, "norm1", groups)), p, "conv1")
    h = _conv(F.silu(_gn(h, p, "norm2", groups)), p, "conv2")
    if p.has("conv_shortcut.weight"):
        x = _conv(x, p, "conv_shortcut", padding=0)
    return x + h


def attention_block(x, p: P, groups=GROUPS):
    b, c, h, w = x.shape
    y = _gn(x, p, "group_norm", groups).view(b, c, h * w).transpose(1, 2)
    q, k, v = linear(y, p, "query"), linear(y, p, "key"), linear(y, p, "value")
    s = torch.baddbmm(torch.empty(b, h * w, h * w, dtype=q.dtype, device=q.device), q, k.transpose(-1, -2), beta=0,
                      alpha=c ** -0.5)
    pr = torch.softmax(s.float(), dim=-1).type(s.dtype)
    o = linear(torch.bmm(pr, v), p, "proj_attn")
    return o.transpose(-1, -2).reshape(b, c, h, w) + x


def mid_block(x, p: P, groups=GROUPS):
    x = resnet2d(x, p.sub("resnets.0"), groups)
    x = attention_block(x, p.sub("attentions.0"), groups)
    return resnet2d(x, p.sub("resnets.1"), groups)


def decode(sd, z, groups=GROUPS):
    """AutoencoderKL.decode (vae.py:575-610): post_quant_conv -> Decoder."""
    p = P(sd)
    x = _conv(z, p, "post_quant_conv", padding=0)
    d = p.sub("decoder")
    x = _conv(x, d, "conv_in")
    x = mid_block(x, d.sub("mid_block"), groups)
    for i in range(d.count("up_blocks")):
        bp = d.sub(f"up_blocks.{i}")
        for j in range(bp.count("resnets")):
            x = resnet2d(x, bp.sub(f"resnets.{j}"), groups)
        if bp.has_prefix("upsamplers"):
            x = F.interpolate(x.float(), scale_factor=2.0, mode="nearest").to(x.dtype)
            x = _conv(x, bp, "upsamplers.0.conv")
    return _conv(F.silu(_gn(x, d, "conv_norm_out", groups)), d, "conv_out")


def encode_moments(sd, x, groups=GROUPS):
    """AutoencoderKL.encode (vae.py:565-573): Encoder -> quant_conv -> (mean, logvar)."""
    p = P(sd)
    e = p.sub("encoder")
    x = _conv(x, e, "conv_in")
    for i in range(e.count("down_blocks")):
        bp = e.sub(f"down_blocks.{i}")
        for j in range(bp.count("resnets")):
            x = resnet2d(x, bp.sub(f"resnets.{j}"), groups)
        if bp.has_prefix("downsamplers"):
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)   # resnet.py:184 asymmetric pad
            x = _conv(x, bp, "downsamplers.0.conv", stride=2, padding=0)
    x = mid_block(x, e.sub("mid_block"), groups)
    x = _conv(F.silu(_gn(x, e, "conv_norm_out", groups)), e, "conv_out")
    return _conv(x, p, "quant_conv", padding=0)


def sample_posterior(moments, noise):
    """DiagonalGaussianDistribution.sample (vae.py:340-361) with the randn draw passed in."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return mean + std * noise.to(moments.dtype)
