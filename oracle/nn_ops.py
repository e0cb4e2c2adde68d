"""Primitive ops of the oracle (plain torch, any float dtype, CPU or CUDA)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class P:
    """A view of a flat state dict under a key prefix: ``P(sd, "down_blocks.0")["resnets.0.norm1.weight"]``."""

    def __init__(self, sd, prefix=""):
        self.sd = sd
        self.prefix = prefix

    def sub(self, name):
        return P(self.sd, f"{self.prefix}{name}.")

    def __getitem__(self, key):
        return self.sd[self.prefix + key]

    def get(self, key, default=None):
        return self.sd.get(self.prefix + key, default)

    def has(self, key):
        return (self.prefix + key) in self.sd

    def has_prefix(self, name):
        pre = f"{self.prefix}{name}."
        return any(k.startswith(pre) for k in self.sd)

    def count(self, name):
        """Number of consecutive integer children ``name.0, name.1, ...``."""
        n = 0
        while self.has_prefix(f"{name}.{n}"):
            n += 1
        return n


def linear(x, p: P, name: str):
    return F.linear(x, p[f"{name}.weight"], p.get(f"{name}.bias"))


def group_norm_frames(x, p: P, name: str, groups: int, eps: float):
    """InflatedGroupNorm: statistics per (batch, frame) image (animatediff/models/resnet.py:9-17).
    x: [b, c, f, h, w]."""
    b, c, f, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    y = F.group_norm(y, groups, p[f"{name}.weight"], p[f"{name}.bias"], eps)
    return y.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def conv2d_frames(x, p: P, name: str, stride=1, padding=1):
    """InflatedConv3d: a 2-D convolution applied to every frame (resnet.py:19-27)."""
    b, c, f, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    y = F.conv2d(y, p[f"{name}.weight"], p.get(f"{name}.bias"), stride=stride, padding=padding)
    return y.reshape(b, f, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def split_heads(t, heads):
    b, n, c = t.shape
    return t.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3)  # [b, h, n, d]


def merge_heads(t):
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


def attention_core(q, k, v, heads, bias=None, math_path=False):
    """softmax(q k^T / sqrt(d) + bias) v with q/k/v [b, n, heads*d].

    ``math_path=False`` is the fused FMHA semantics (xformers / SDPA: fp32 softmax inside).
    ``math_path=True`` follows the reference's baddbmm -> softmax -> bmm in the tensor dtype
    (diffusers/models/attention_processor.py:562-591) -- identical in fp32."""
    qh, kh, vh = split_heads(q, heads), split_heads(k, heads), split_heads(v, heads)
    scale = qh.shape[-1] ** -0.5
    if math_path:
        s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
        if bias is not None:
            s = s + bias
        pr = s.softmax(dim=-1).to(vh.dtype)
        o = torch.matmul(pr, vh)
    else:
        o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=bias)
    return merge_heads(o)


def geglu_ff(x, p: P):
    """FeedForward with GEGLU (diffusers/models/attention_lora.py:493-547, activations.py:93-122):
    net.0.proj: C -> 8C, value * gelu_erf(gate), net.2: 4C -> C."""
    h = linear(x, p, "net.0.proj")
    val, gate = h.chunk(2, dim=-1)
    return linear(val * F.gelu(gate), p, "net.2")


def timestep_embedding(t, dim, flip_sin_to_cos=True, freq_shift=0.0, max_period=10000):
    """Timesteps / get_timestep_embedding (diffusers/models/embeddings.py:26-66). Returns fp32."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device)
    exponent = exponent / (half - freq_shift)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def timestep_mlp(x, p: P):
    """TimestepEmbedding: linear_1 -> SiLU -> linear_2 (embeddings.py:190-236)."""
    return linear(F.silu(linear(x, p, "linear_1")), p, "linear_2")


def sinusoid_table(max_len, d_model, device=None):
    """PositionalEncoding buffer (animatediff/models/motion_module.py:262-280)."""
    position = torch.arange(max_len, device=device).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, device=device) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model, device=device)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe
