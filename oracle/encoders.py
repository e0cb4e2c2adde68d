"""Oracle of SURVEY.md section 8(f) row 2: the two conditioning encoders the pipeline runs once per clip.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain functional torch, state dict in, tensor out.

Neither encoder's source lives under /root/reference -- both are third-party packages the reference imports:

* ``transformers.CLIPTextModel`` (requirements.txt:8 ``transformers``, unpinned; built at inference_dual_p2e.py:387, called
  at animatediff/pipelines/pipeline_animation_inference_dual.py:236-240,:283-287 -> ``[0]`` = last_hidden_state).
  Restated here from the published model (pre-LN transformer, causal mask, erf-GELU or quick-GELU MLP, final LayerNorm)
  and PINNED against the transformers 5.5.0 installed in this image (tools/make_golden_encoders.py ->
  tests/golden/encoders.pt; tests/test_encoders_oracle.py also compares live when transformers imports).
* ``segment_anything`` (requirements.txt:11, unpinned, NOT installed here): ``sam_model_registry["vit_b"]`` ->
  ``Sam.image_encoder`` = ``ImageEncoderViT`` (patch 16, 768 wide, 12 blocks, 12 heads, 14 x 14 windows with global
  attention in blocks 2/5/8/11, decomposed relative position bias, 256-channel neck), driven through
  ``SamPredictor.set_torch_image`` (pipeline...dual.py:685-690,:708-713).  Restated from the published algorithm
  (segment_anything/modeling/image_encoder.py of release 1.0) with ITS parameter names, and pinned against the only
  other implementation available offline: transformers' ``SamVisionEncoder`` port (same fixture file; ``hf_sam_keys``
  maps the names).  The package itself being absent, parity with segment_anything proper stays "unpinned".
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# CLIP text encoder
# ------------------------------------------------------------------------------------------------
def _act(x, name):
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if name in ("gelu", "gelu_python"):
        return F.gelu(x)
    raise ValueError(name)


def clip_text_forward(sd, input_ids, heads: int, eps: float = 1e-5, act: str = "gelu", prefix: str = "text_model."):
    """``CLIPTextModel(input_ids)[0]``: token + position embedding, pre-LN blocks with a causal mask, final LayerNorm."""
    g = lambda k: sd[prefix + k]
    x = g("embeddings.token_embedding.weight")[input_ids] + g("embeddings.position_embedding.weight")[: input_ids.shape[1]]
    b, n, c = x.shape
    hd = c // heads
    causal = torch.full((n, n), float("-inf"), dtype=x.dtype, device=x.device).triu(1)
    i = 0
    while prefix + f"encoder.layers.{i}.layer_norm1.weight" in sd:
        L = f"encoder.layers.{i}."
        h = F.layer_norm(x, (c,), g(L + "layer_norm1.weight"), g(L + "layer_norm1.bias"), eps)
        q = F.linear(h, g(L + "self_attn.q_proj.weight"), g(L + "self_attn.q_proj.bias"))
        k = F.linear(h, g(L + "self_attn.k_proj.weight"), g(L + "self_attn.k_proj.bias"))
        v = F.linear(h, g(L + "self_attn.v_proj.weight"), g(L + "self_attn.v_proj.bias"))
        sp = lambda t: t.view(b, n, heads, hd).transpose(1, 2)
        a = (sp(q) @ sp(k).transpose(-1, -2)) * hd ** -0.5 + causal
        o = (a.softmax(-1) @ sp(v)).transpose(1, 2).reshape(b, n, c)
        x = x + F.linear(o, g(L + "self_attn.out_proj.weight"), g(L + "self_attn.out_proj.bias"))
        h = F.layer_norm(x, (c,), g(L + "layer_norm2.weight"), g(L + "layer_norm2.bias"), eps)
        h = _act(F.linear(h, g(L + "mlp.fc1.weight"), g(L + "mlp.fc1.bias")), act)
        x = x + F.linear(h, g(L + "mlp.fc2.weight"), g(L + "mlp.fc2.bias"))
        i += 1
    return F.layer_norm(x, (c,), g("final_layer_norm.weight"), g("final_layer_norm.bias"), eps)


# ------------------------------------------------------------------------------------------------
# SAM image encoder (segment_anything ImageEncoderViT)
# ------------------------------------------------------------------------------------------------
def sam_preprocess(x, pixel_mean, pixel_std, img_size: int):
    """``Sam.preprocess``: normalise, zero-pad bottom / right to the square input."""
    x = (x - pixel_mean.view(-1, 1, 1)) / pixel_std.view(-1, 1, 1)
    return F.pad(x, (0, img_size - x.shape[-1], 0, img_size - x.shape[-2]))


def _rel_pos(q_size: int, k_size: int, rel_pos):
    """``get_rel_pos``: table [L, c] -> [q_size, k_size, c] (linear interpolation when L != 2 max(q, k) - 1)."""
    dist = 2 * max(q_size, k_size) - 1
    if rel_pos.shape[0] != dist:
        rel_pos = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=dist, mode="linear")
        rel_pos = rel_pos.reshape(-1, dist).permute(1, 0)
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    idx = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rel_pos[idx.long()]


def _sam_attention(x, p, heads: int):
    """``Attention.forward`` on [B, H, W, C]: logits = (q * scale) k^T + q . Rh[qh, kh] + q . Rw[qw, kw] (the relative
    position terms use the UNSCALED q)."""
    b, hh, ww, c = x.shape
    hd = c // heads
    qkv = F.linear(x, p("attn.qkv.weight"), p("attn.qkv.bias")).reshape(b, hh * ww, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, b * heads, hh * ww, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    rh, rw = _rel_pos(hh, hh, p("attn.rel_pos_h")), _rel_pos(ww, ww, p("attn.rel_pos_w"))
    rq = q.reshape(b * heads, hh, ww, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, rw)
    attn = (attn.view(b * heads, hh, ww, hh, ww) + rel_h[..., :, None] + rel_w[..., None, :]).view(b * heads, hh * ww, hh * ww)
    o = (attn.softmax(-1) @ v).view(b, heads, hh, ww, hd).permute(0, 2, 3, 1, 4).reshape(b, hh, ww, c)
    return F.linear(o, p("attn.proj.weight"), p("attn.proj.bias"))


def window_partition(x, ws: int):
    b, h, w, c = x.shape
    ph, pw = (ws - h % ws) % ws, (ws - w % ws) % ws
    x = F.pad(x, (0, 0, 0, pw, 0, ph))
    hp, wp = h + ph, w + pw
    x = x.view(b, hp // ws, ws, wp // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, c)
    return x, (hp, wp)


def window_unpartition(win, ws: int, pad_hw, hw):
    hp, wp = pad_hw
    h, w = hw
    b = win.shape[0] // (hp * wp // ws // ws)
    x = win.view(b, hp // ws, wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(b, hp, wp, -1)
    return x[:, :h, :w, :]


def _ln2d(x, w, b, eps):
    """``LayerNorm2d``: over the channel dim of NCHW."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


def sam_image_encoder_forward(sd, x, heads: int, window_size: int, global_attn_indexes, eps: float = 1e-6, prefix: str = ""):
    """``ImageEncoderViT.forward``: x [B, 3, S, S] (already preprocessed) -> [B, out_chans, S/patch, S/patch]."""
    g = lambda k: sd[prefix + k]
    w = g("patch_embed.proj.weight")
    patch = w.shape[-1]
    x = F.conv2d(x, w, g("patch_embed.proj.bias"), stride=patch).permute(0, 2, 3, 1)
    if prefix + "pos_embed" in sd:
        x = x + g("pos_embed")
    c = x.shape[-1]
    i = 0
    while prefix + f"blocks.{i}.norm1.weight" in sd:
        p = lambda k, i=i: g(f"blocks.{i}." + k)
        ws = 0 if i in global_attn_indexes else window_size
        short = x
        h = F.layer_norm(x, (c,), p("norm1.weight"), p("norm1.bias"), eps)
        if ws > 0:
            hw = h.shape[1:3]
            h, pad_hw = window_partition(h, ws)
        h = _sam_attention(h, p, heads)
        if ws > 0:
            h = window_unpartition(h, ws, pad_hw, hw)
        x = short + h
        h = F.layer_norm(x, (c,), p("norm2.weight"), p("norm2.bias"), eps)
        x = x + F.linear(F.gelu(F.linear(h, p("mlp.lin1.weight"), p("mlp.lin1.bias"))), p("mlp.lin2.weight"), p("mlp.lin2.bias"))
        i += 1
    x = F.conv2d(x.permute(0, 3, 1, 2), g("neck.0.weight"))
    x = _ln2d(x, g("neck.1.weight"), g("neck.1.bias"), eps)
    x = F.conv2d(x, g("neck.2.weight"), padding=1)
    return _ln2d(x, g("neck.3.weight"), g("neck.3.bias"), eps)


def hf_sam_keys(sd):
    """transformers ``SamVisionEncoder`` parameter names -> segment_anything ``ImageEncoderViT`` names."""
    out = {}
    for k, v in sd.items():
        if k.startswith("neck."):
            k = (k.replace("neck.conv1.", "neck.0.").replace("neck.layer_norm1.", "neck.1.")
                 .replace("neck.conv2.", "neck.2.").replace("neck.layer_norm2.", "neck.3."))
        else:
            k = (k.replace("layers.", "blocks.").replace(".layer_norm1.", ".norm1.").replace(".layer_norm2.", ".norm2.")
                 .replace("patch_embed.projection.", "patch_embed.proj."))
        out[k] = v
    return out


def sa_to_hf_sam_keys(sd):
    """The inverse of :func:`hf_sam_keys` (used by the fixture generator to load synthetic weights into the HF port)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("neck."):
            k = (k.replace("neck.0.", "neck.conv1.").replace("neck.1.", "neck.layer_norm1.")
                 .replace("neck.2.", "neck.conv2.").replace("neck.3.", "neck.layer_norm2."))
        else:
            k = (k.replace("blocks.", "layers.").replace(".norm1.", ".layer_norm1.").replace(".norm2.", ".layer_norm2.")
                 .replace("patch_embed.proj.", "patch_embed.projection."))
        out[k] = v
    return out


def sam_shapes(embed=128, depth=3, heads=2, patch=8, img=112, window=3, global_idx=(1,), out_chans=32, mlp_ratio=4):
    """{key: shape} of an ImageEncoderViT with the given hyper-parameters (segment_anything names)."""
    hd, grid = embed // heads, img // patch
    s = {"pos_embed": (1, grid, grid, embed), "patch_embed.proj.weight": (embed, 3, patch, patch), "patch_embed.proj.bias": (embed,)}
    for i in range(depth):
        S = grid if i in global_idx else window
        b = f"blocks.{i}."
        s.update({b + "norm1.weight": (embed,), b + "norm1.bias": (embed,), b + "norm2.weight": (embed,), b + "norm2.bias": (embed,),
                  b + "attn.qkv.weight": (3 * embed, embed), b + "attn.qkv.bias": (3 * embed,),
                  b + "attn.proj.weight": (embed, embed), b + "attn.proj.bias": (embed,),
                  b + "attn.rel_pos_h": (2 * S - 1, hd), b + "attn.rel_pos_w": (2 * S - 1, hd),
                  b + "mlp.lin1.weight": (mlp_ratio * embed, embed), b + "mlp.lin1.bias": (mlp_ratio * embed,),
                  b + "mlp.lin2.weight": (embed, mlp_ratio * embed), b + "mlp.lin2.bias": (embed,)})
    s.update({"neck.0.weight": (out_chans, embed, 1, 1), "neck.1.weight": (out_chans,), "neck.1.bias": (out_chans,),
              "neck.2.weight": (out_chans, out_chans, 3, 3), "neck.3.weight": (out_chans,), "neck.3.bias": (out_chans,)})
    return s


def clip_shapes(hidden=128, inter=256, layers=2, vocab=120, max_pos=77):
    s = {"text_model.embeddings.token_embedding.weight": (vocab, hidden),
         "text_model.embeddings.position_embedding.weight": (max_pos, hidden),
         "text_model.final_layer_norm.weight": (hidden,), "text_model.final_layer_norm.bias": (hidden,)}
    for i in range(layers):
        L = f"text_model.encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[L + f"self_attn.{n}.weight"] = (hidden, hidden)
            s[L + f"self_attn.{n}.bias"] = (hidden,)
        s.update({L + "layer_norm1.weight": (hidden,), L + "layer_norm1.bias": (hidden,), L + "layer_norm2.weight": (hidden,),
                  L + "layer_norm2.bias": (hidden,), L + "mlp.fc1.weight": (inter, hidden), L + "mlp.fc1.bias": (inter,),
                  L + "mlp.fc2.weight": (hidden, inter), L + "mlp.fc2.bias": (hidden,)})
    return s
