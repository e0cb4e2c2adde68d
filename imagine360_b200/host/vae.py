"""Host-side mirror of ``diffusers.AutoencoderKL`` (the vendored 0.11-era tree: diffusers/models/vae.py:501-610) with the
SD-2.1 VAE topology; same constructor kwargs, parameter names and ``encode(x, *).latent_dist.sample()`` /
``decode(z).sample`` surface.  All convolutions / norms / the mid-block attention run on the native kernels,
channels-last; frames are batched (the reference decodes frame by frame with batch 1,
pipeline_animation_inference_dual.py:306-307, which changes nothing numerically: no op couples samples).
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import forward as Fw
from .forward import tokens
from .unet3d import BF16, cached, conv_w, lin_w, pad8, upsample_conv_w


class ResnetBlock2D(nn.Module):
    """diffusers/models/resnet.py:367-496 (temb_channels=None)"""

    def __init__(self, cin, cout, groups, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None
        self.output_scale_factor = 1.0


class AttentionBlock(nn.Module):
    """diffusers/models/attention.py:247-380 (single head)"""

    def __init__(self, channels, groups, eps=1e-6):
        super().__init__()
        self.channels = channels
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.query, self.key, self.value = (nn.Linear(channels, channels) for _ in range(3))
        self.proj_attn = nn.Linear(channels, channels)


class _Sampler(nn.Module):
    def __init__(self, ch, stride):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=stride, padding=1 if stride == 1 else 0)


class _Blk(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.downsamplers = None
        self.upsamplers = None


class _Mid(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, groups), ResnetBlock2D(ch, ch, groups)])
        self.attentions = nn.ModuleList([AttentionBlock(ch, groups)])


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        ch = block_out_channels[0]
        for i, co in enumerate(block_out_channels):
            blk = _Blk()
            for j in range(layers_per_block):
                blk.resnets.append(ResnetBlock2D(ch if j == 0 else co, co, groups))
            ch = co
            if i != len(block_out_channels) - 1:
                blk.downsamplers = nn.ModuleList([_Sampler(co, 2)])
            self.down_blocks.append(blk)
        self.mid_block = _Mid(ch, groups)
        self.conv_norm_out = nn.GroupNorm(groups, ch, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch, 2 * out_channels, 3, padding=1)


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        ch = rev[0]
        for i, co in enumerate(rev):
            blk = _Blk()
            for j in range(layers_per_block + 1):
                blk.resnets.append(ResnetBlock2D(ch if j == 0 else co, co, groups))
            ch = co
            if i != len(rev) - 1:
                blk.upsamplers = nn.ModuleList([_Sampler(co, 1)])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(groups, ch, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch, out_channels, 3, padding=1)


# ------------------------------------------------------------------------------------------------------
def _to_nhwc(x):
    n, c, h, w = x.shape
    y = x.permute(0, 2, 3, 1).to(BF16)
    if pad8(c) != c:
        y = F.pad(y, (0, pad8(c) - c))
    return y.contiguous()


# GroupNorm statistics from the producing conv's epilogue (group sizes 4 / 8 / 16 = the VAE's widths with 32 groups):
# I360_VAE_GN_FUSED=0 restores the separate statistics pass (A/B runs)
GN_FUSED = os.environ.get("I360_VAE_GN_FUSED", "1") not in ("", "0")


def _gn_ok(c, groups):
    return GN_FUSED and groups == 32 and c // groups in (4, 8, 16) and c % 32 == 0


def _conv_stats(x, w, b, cout, groups):
    """conv3x3 (+ bias) -> (y, GroupNorm statistics of y | None)"""
    if _gn_ok(cout, groups):
        return ops.conv3x3(x, w, bias=b, gn_groups=groups)
    return ops.conv3x3(x, w, bias=b), None


def _resnet(x, r, groups, x_stats=None, want_stats=False):
    """-> (y, GroupNorm statistics of y | None).  ``x_stats``: statistics of ``x`` from the conv that produced it."""
    h = ops.groupnorm(x, r.norm1.weight, r.norm1.bias, groups, r.norm1.eps, True, stats=x_stats)
    w1, b1 = conv_w(r.conv1)
    co = r.conv1.out_channels
    if _gn_ok(co, groups):
        h, st = ops.conv3x3(h, w1, bias=b1, gn_groups=groups)
    else:
        h, st = ops.conv3x3(h, w1, bias=b1), None
    h = ops.groupnorm(h, r.norm2.weight, r.norm2.bias, groups, r.norm2.eps, True, stats=st)
    gg = groups if (want_stats and _gn_ok(co, groups)) else None
    if r.conv_shortcut is not None:
        w2, b2 = conv_w(r.conv2, shortcut=r.conv_shortcut)
        y = ops.conv3x3(h, w2, bias=b2, x2=x, gn_groups=gg)
    else:
        w2, b2 = conv_w(r.conv2)
        y = ops.conv3x3(h, w2, bias=b2, resid=x, gn_groups=gg)
    return y if gg is not None else (y, None)


def _attention(x, a, groups):
    """Single-head attention with head_dim = C (512): scores are materialised per image
    (baddbmm -> fp32 softmax -> bmm, attention.py:353-367) with the tensor-core GEMM."""
    n, h, w, c = x.shape
    hn = ops.groupnorm(x, a.group_norm.weight, a.group_norm.bias, groups, a.group_norm.eps, False)
    t = tokens(hn)
    wq = cached(a, "qkv", [a.query.weight, a.key.weight, a.value.weight, a.query.bias, a.key.bias, a.value.bias],
                lambda: (torch.cat([a.query.weight, a.key.weight, a.value.weight], 0).to(BF16).contiguous(),
                         torch.cat([a.query.bias, a.key.bias, a.value.bias], 0).to(BF16).contiguous()))
    qkv = ops.gemm(t, wq[0], bias=wq[1]).view(n, h * w, 3 * c)
    out = torch.empty((n, h * w, c), dtype=BF16, device=x.device)
    for i in range(n):
        q, k, v = qkv[i, :, :c], qkv[i, :, c:2 * c], qkv[i, :, 2 * c:]
        s = ops.gemm(q, k, out_scale=c ** -0.5)                       # [N, N] bf16, like baddbmm(alpha=scale)
        p = ops.softmax_rows(s)
        ops.gemm(p, v.t().contiguous(), out=out[i])                   # P @ V  (V^T is the [C, N] "weight")
    wo, bo = lin_w(a.proj_attn)
    return ops.gemm(out.view(-1, c), wo, bias=bo, resid=tokens(x)).view(n, h, w, c)


def _mid(x, m, groups, x_stats=None, want_stats=False):
    x, _ = _resnet(x, m.resnets[0], groups, x_stats)
    x = _attention(x, m.attentions[0], groups)
    return _resnet(x, m.resnets[1], groups, None, want_stats)


class DiagonalGaussianDistribution:
    """diffusers/models/vae.py:340-361"""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device)
        return self.mean + self.std * noise.to(self.parameters.dtype)

    def mode(self):
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",), up_block_types=("UpDecoderBlock2D",),
                 block_out_channels=(64,), layers_per_block=1, act_fn="silu", latent_channels=4, norm_num_groups=32,
                 sample_size=32, **extra):
        super().__init__()
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                                      layers_per_block=layers_per_block, latent_channels=latent_channels,
                                      norm_num_groups=norm_num_groups, sample_size=sample_size, **extra)
        self.groups = norm_num_groups
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.use_slicing = False

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    @property
    def device(self):
        return self.quant_conv.weight.device

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **_):
        p = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(p, "config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        m = cls(**cfg)
        for name in ("diffusion_pytorch_model.bin", "diffusion_pytorch_model.safetensors"):
            wp = os.path.join(p, name)
            if os.path.isfile(wp):
                if name.endswith(".bin"):
                    sd = torch.load(wp, map_location="cpu")
                else:
                    from safetensors.torch import load_file
                    sd = load_file(wp)
                m.load_state_dict(sd, strict=False)
                break
        return m

    def _conv1x1(self, x, conv):
        """quant_conv / post_quant_conv on NHWC tokens (channels padded to 8)."""
        co, ci = conv.weight.shape[:2]
        w = cached(conv, "w1", [conv.weight, conv.bias],
                   lambda: (F.pad(conv.weight.to(BF16).reshape(co, ci), (0, pad8(ci) - ci, 0, pad8(co) - co)).contiguous(),
                            F.pad(conv.bias.to(BF16), (0, pad8(co) - co)).contiguous()))
        n, h, ww, c = x.shape
        return ops.gemm(tokens(x), w[0], bias=w[1]).view(n, h, ww, -1)

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        with ops.on_device(x.device):
            return self._encode(x, return_dict)

    def _encode(self, x, return_dict=True):
        """vae.py:565-573 / Encoder.forward :128-144.  (The pipeline passes the chunk length as the 2nd positional,
        which lands in ``return_dict`` -- any truthy value keeps the output object form; SURVEY.md trap 10.)"""
        e, g = self.encoder, self.groups
        y = _to_nhwc(x)
        w, b = conv_w(e.conv_in, cin_pad=y.shape[-1])
        y, st = _conv_stats(y, w, b, e.conv_in.out_channels, g)
        for blk in e.down_blocks:
            for r in blk.resnets:
                y, st = _resnet(y, r, g, st, True)
            if blk.downsamplers is not None:
                st = None                # the stride-2 conv has no statistics epilogue: the next norm1 runs its own pass
                n, h, ww, c = y.shape
                wd, bd = conv_w(blk.downsamplers[0].conv)
                if Fw.S2_IM2COL:
                    y = ops.gemm(ops.im2col_s2(y, False, pad_lo=0), wd, bias=bd).view(n, h // 2, ww // 2, -1)
                else:                    # asymmetric F.pad(0, 1, 0, 1) + stride-2 conv: TMA boxes with traversal stride 2
                    y = ops.conv3x3_s2(y, wd, bd, pad_lo=0)
        y, st = _mid(y, e.mid_block, g, st, True)
        y = ops.groupnorm(y, e.conv_norm_out.weight, e.conv_norm_out.bias, g, e.conv_norm_out.eps, True, stats=st)
        w, b = conv_w(e.conv_out, cout_pad=pad8(e.conv_out.out_channels))
        y = ops.conv3x3(y, w, bias=b)
        y = self._conv1x1(y, self.quant_conv)
        moments = y[..., : self.quant_conv.out_channels].permute(0, 3, 1, 2).contiguous()
        post = DiagonalGaussianDistribution(moments)
        return SimpleNamespace(latent_dist=post) if return_dict else (post,)

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        with ops.on_device(z.device):
            return self._decode(z, return_dict)

    def _decode(self, z, return_dict=True):
        """vae.py:575-610 / Decoder.forward :208-224"""
        d, g = self.decoder, self.groups
        y = self._conv1x1(_to_nhwc(z), self.post_quant_conv)
        w, b = conv_w(d.conv_in, cin_pad=y.shape[-1])
        y, st = _conv_stats(y, w, b, d.conv_in.out_channels, g)
        y, st = _mid(y, d.mid_block, g, st, True)
        for blk in d.up_blocks:
            for r in blk.resnets:
                y, st = _resnet(y, r, g, st, True)
            if blk.upsamplers is not None:
                co = blk.upsamplers[0].conv.out_channels
                if Fw.SUBPIXEL:          # nearest x2 + conv3x3 as four 2x2-tap convolutions of the low-resolution tensor
                    wu, bu = upsample_conv_w(blk.upsamplers[0].conv)
                    if _gn_ok(co, g):
                        y, st = ops.conv_upsample2x(y, wu, bu, gn_groups=g)
                    else:
                        y, st = ops.conv_upsample2x(y, wu, bu), None
                else:
                    wu, bu = conv_w(blk.upsamplers[0].conv)
                    y, st = _conv_stats(ops.upsample2x(y), wu, bu, co, g)
        y = ops.groupnorm(y, d.conv_norm_out.weight, d.conv_norm_out.bias, g, d.conv_norm_out.eps, True, stats=st)
        w, b = conv_w(d.conv_out, cout_pad=pad8(d.conv_out.out_channels))
        y = ops.conv3x3(y, w, bias=b)
        out = y[..., : d.conv_out.out_channels].permute(0, 3, 1, 2).contiguous()
        return SimpleNamespace(sample=out) if return_dict else (out,)
