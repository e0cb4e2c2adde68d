"""Conditioning encoders run once per clip on the same sm_100a kernels as the denoiser.

The reference builds both from third-party packages and hands the module instances to ``AnimationPipeline``
(inference_dual_p2e.py:369-370,:387,:466-476):

* ``transformers.CLIPTextModel`` -> ``_encode_prompt`` (pipeline_animation_inference_dual.py:236-240,:283-287, ``[0]``),
* ``segment_anything`` ``Sam`` -> ``SamPredictor.set_torch_image`` -> ``Sam.image_encoder`` (:685-690,:708-713).

The boundary here is therefore the module instance: :class:`ClipTextNative` / :class:`SamImageEncoderNative` wrap the
caller's module, read ITS parameters (by the packages' own names) and hyper-parameters, and evaluate the same forward
as C-ABI calls -- LayerNorm, tcgen05 GEMMs with fused bias / GELU / residual epilogues, the flash attention kernel
(causal mask as a dense bias for CLIP; SAM's decomposed relative position term as a per-(window, head) bias built by
``i360_relpos_bias_bf16``), the implicit-GEMM 3x3 conv of SAM's neck.  bf16 with fp32 accumulation; the reference runs
both encoders in fp32 and casts their outputs to bf16 (pipeline...dual.py:695,:716 ``.to(dtype=latents_dtype)``; text
embeddings are consumed by bf16 UNets).  There is no fallback: a CUDA input always takes this path; an unknown module
layout raises.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import ops
from .unet3d import BF16, cached


def _bf(t):
    return t.detach().to(BF16).contiguous()


class _Out(tuple):
    """What ``CLIPTextModel.forward`` returns as far as the pipeline looks: ``out[0]`` / ``out.last_hidden_state``."""

    @property
    def last_hidden_state(self):
        return self[0]


# ------------------------------------------------------------------------------------------------------
# CLIP text encoder
# ------------------------------------------------------------------------------------------------------
class ClipTextNative:
    """Native forward of a ``transformers.CLIPTextModel`` (``text_model.embeddings`` / ``encoder.layers.N`` /
    ``final_layer_norm``), called like the module: ``enc(input_ids, attention_mask=None)[0]``."""

    def __init__(self, module):
        tm = getattr(module, "text_model", None)
        if tm is None or not hasattr(tm, "encoder") or not hasattr(tm, "embeddings"):
            raise TypeError("imagine360_b200: ClipTextNative needs a transformers CLIPTextModel (text_model.embeddings / .encoder)")
        self.module = module
        self.config = module.config
        cfg = module.config
        self.heads = int(cfg.num_attention_heads)
        self.eps = float(getattr(cfg, "layer_norm_eps", 1e-5))
        act = getattr(cfg, "hidden_act", "gelu")
        if act not in ("gelu", "quick_gelu"):
            raise NotImplementedError(f"imagine360_b200: CLIP hidden_act {act!r}")
        self.act = act

    @staticmethod
    def supports(module) -> bool:
        tm = getattr(module, "text_model", None)
        cfg = getattr(module, "config", None)
        return (tm is not None and hasattr(tm, "encoder") and cfg is not None and
                cfg.hidden_size % cfg.num_attention_heads == 0 and cfg.hidden_size // cfg.num_attention_heads == 64)

    @property
    def device(self):
        return next(self.module.parameters()).device

    def _layer(self, layer):
        sa, mlp = layer.self_attn, layer.mlp
        params = [sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight, sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias,
                  sa.out_proj.weight, sa.out_proj.bias, mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias,
                  layer.layer_norm1.weight, layer.layer_norm1.bias, layer.layer_norm2.weight, layer.layer_norm2.bias]

        def build():
            return dict(wqkv=_bf(torch.cat([sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight], 0)),
                        bqkv=_bf(torch.cat([sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias], 0)),
                        wo=_bf(sa.out_proj.weight), bo=_bf(sa.out_proj.bias), w1=_bf(mlp.fc1.weight), b1=_bf(mlp.fc1.bias),
                        w2=_bf(mlp.fc2.weight), b2=_bf(mlp.fc2.bias), g1=_bf(layer.layer_norm1.weight),
                        be1=_bf(layer.layer_norm1.bias), g2=_bf(layer.layer_norm2.weight), be2=_bf(layer.layer_norm2.bias))
        return cached(layer, "clip_layer", params, build)

    @torch.no_grad()
    def __call__(self, input_ids, attention_mask=None, **_):
        if attention_mask is not None:
            raise NotImplementedError("imagine360_b200: CLIP text encoder with an attention mask (SD-2.1's config has "
                                      "no use_attention_mask; pipeline...dual.py:230-233 passes None)")
        tm = self.module.text_model
        dev = tm.embeddings.token_embedding.weight.device
        with ops.on_device(dev):
            ids = input_ids.to(dev)
            b, n = ids.shape
            emb = tm.embeddings
            tabs = cached(emb, "clip_emb", [emb.token_embedding.weight, emb.position_embedding.weight],
                          lambda: (_bf(emb.token_embedding.weight), _bf(emb.position_embedding.weight)))
            # the embedding lookup is a row gather of 2 x 77 rows: torch indexing (plumbing); the sum is done in fp32 like
            # the reference (fp32 module) and rounded once
            x = (tabs[0][ids].float() + tabs[1][:n].float()).to(BF16).view(b * n, -1)
            c = x.shape[1]
            hd = c // self.heads
            causal = cached(emb, f"clip_causal_{n}_{dev}", [], lambda: torch.full((n, n), float("-inf"), device=dev).triu(1).to(BF16))
            for layer in tm.encoder.layers:
                w = self._layer(layer)
                h = ops.layernorm(x, w["g1"], w["be1"], self.eps)
                qkv = ops.gemm(h, w["wqkv"], bias=w["bqkv"])
                o = torch.empty((b * n, c), dtype=BF16, device=dev)
                ops.attention(ops.seq_view(qkv, b, n, 0), ops.seq_view(qkv, b, n, c), ops.seq_view(qkv, b, n, 2 * c),
                              ops.seq_view(o, b, n, 0), self.heads, hd, b, bias=causal)
                x = ops.gemm(o, w["wo"], bias=w["bo"], resid=x)
                h = ops.layernorm(x, w["g2"], w["be2"], self.eps)
                if self.act == "gelu":
                    h = ops.gemm(h, w["w1"], bias=w["b1"], act=ops.ACT_GELU)
                else:
                    h = ops.gemm(h, w["w1"], bias=w["b1"]).float()
                    h = (h * torch.sigmoid(1.702 * h)).to(BF16)          # OpenAI CLIP towers only (not SD-2.1's)
                x = ops.gemm(h, w["w2"], bias=w["b2"], resid=x)
            fl = tm.final_layer_norm
            g, be = cached(fl, "clip_final", [fl.weight, fl.bias], lambda: (_bf(fl.weight), _bf(fl.bias)))
            y = ops.layernorm(x, g, be, self.eps).view(b, n, c)
        return _Out((y.to(tm.embeddings.token_embedding.weight.dtype),))


# ------------------------------------------------------------------------------------------------------
# SAM ViT image encoder
# ------------------------------------------------------------------------------------------------------
class SamImageEncoderNative:
    """Native forward of segment_anything's ``ImageEncoderViT`` (``patch_embed.proj``, ``pos_embed``, ``blocks.N`` with
    ``norm1 / attn.{qkv, proj, rel_pos_h, rel_pos_w} / norm2 / mlp.{lin1, lin2}`` and ``window_size``, ``neck.0..3``):
    ``enc(x)`` with x [B, 3, S, S] preprocessed -> [B, out_chans, S / patch, S / patch]."""

    def __init__(self, module):
        if not self.supports(module):
            raise TypeError("imagine360_b200: SamImageEncoderNative needs a segment_anything ImageEncoderViT "
                            "(patch_embed.proj / blocks[i].attn.qkv + rel_pos_h / neck) with 64-wide heads")
        self.module = module
        self.img_size = int(module.img_size)

    @staticmethod
    def supports(module) -> bool:
        try:
            blk = module.blocks[0]
            heads = int(blk.attn.num_heads)
            dim = blk.attn.qkv.weight.shape[1]
            return (hasattr(module, "patch_embed") and hasattr(module, "neck") and hasattr(blk.attn, "rel_pos_h") and
                    hasattr(blk, "window_size") and dim // heads == 64 and dim % heads == 0)
        except (AttributeError, IndexError, TypeError):
            return False

    def _block(self, blk):
        a, m = blk.attn, blk.mlp
        params = [a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, a.rel_pos_h, a.rel_pos_w, m.lin1.weight, m.lin1.bias,
                  m.lin2.weight, m.lin2.bias, blk.norm1.weight, blk.norm1.bias, blk.norm2.weight, blk.norm2.bias]

        def build():
            return dict(wqkv=_bf(a.qkv.weight), bqkv=_bf(a.qkv.bias), wo=_bf(a.proj.weight), bo=_bf(a.proj.bias),
                        rh=a.rel_pos_h.detach(), rw=a.rel_pos_w.detach(), w1=_bf(m.lin1.weight), b1=_bf(m.lin1.bias),
                        w2=_bf(m.lin2.weight), b2=_bf(m.lin2.bias), g1=_bf(blk.norm1.weight), be1=_bf(blk.norm1.bias),
                        g2=_bf(blk.norm2.weight), be2=_bf(blk.norm2.bias))
        return cached(blk, "sam_block", params, build)

    @staticmethod
    def _rel_table(tab, S: int):
        """get_rel_pos's table for q_size == k_size == S: [2S - 1, hd] (linear interpolation when the stored length
        differs, as the published code does)."""
        L = 2 * S - 1
        if tab.shape[0] != L:
            tab = F.interpolate(tab.float().reshape(1, tab.shape[0], -1).permute(0, 2, 1), size=L, mode="linear")
            tab = tab.reshape(-1, L).permute(1, 0)
        return tab.to(BF16).contiguous()

    def _attention(self, h, w, items: int, S: int, heads: int):
        """h: [items * S*S, C] normalised tokens of ``items`` sequences of S x S tokens -> attention output (before proj)."""
        c = h.shape[1]
        hd = c // heads
        qkv = ops.gemm(h, w["wqkv"], bias=w["bqkv"])                # columns [q | k | v], heads inside (reshape(.., 3, heads, hd))
        bias = ops.relpos_bias(qkv, 0, items, heads, hd, S, self._rel_table(w["rh"], S), self._rel_table(w["rw"], S))
        o = torch.empty((items * S * S, c), dtype=BF16, device=h.device)
        n = S * S
        ops.attention_item_bias(ops.seq_view(qkv, items, n, 0), ops.seq_view(qkv, items, n, c), ops.seq_view(qkv, items, n, 2 * c),
                                ops.seq_view(o, items, n, 0), heads, hd, items, bias)
        return o

    @torch.no_grad()
    def __call__(self, x):
        mod = self.module
        pw = mod.patch_embed.proj.weight
        dev = pw.device
        with ops.on_device(dev):
            b, cin, hh, ww = x.shape
            p = pw.shape[-1]
            gh, gw = hh // p, ww // p
            c = pw.shape[0]
            # patch embedding = GEMM over non-overlapping patches (the unfold is a permute copy: plumbing), + pos_embed
            wpe, bpe = cached(mod.patch_embed, "sam_pe", [pw, mod.patch_embed.proj.bias],
                              lambda: (_bf(pw.reshape(c, -1)), _bf(mod.patch_embed.proj.bias)))
            cols = x.to(BF16).view(b, cin, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(b * gh * gw, cin * p * p).contiguous()
            pos = None
            if getattr(mod, "pos_embed", None) is not None:
                pos = cached(mod, f"sam_pos_{b}", [mod.pos_embed],
                             lambda: _bf(mod.pos_embed).view(1, gh * gw, c).expand(b, -1, -1).reshape(b * gh * gw, c).contiguous())
            xt = ops.gemm(cols, wpe, bias=bpe, resid=pos)               # [(b gh gw), c]
            for blk in mod.blocks:
                w = self._block(blk)
                heads, ws, eps = int(blk.attn.num_heads), int(blk.window_size), float(blk.norm1.eps)
                h = ops.layernorm(xt, w["g1"], w["be1"], eps)
                if ws > 0:
                    # window_partition: zero-pad the NORMALISED tokens to a multiple of the window, windows become sequences
                    ph, pwd = (ws - gh % ws) % ws, (ws - gw % ws) % ws
                    hp, wp = gh + ph, gw + pwd
                    hw = F.pad(h.view(b, gh, gw, c), (0, 0, 0, pwd, 0, ph))
                    hw = hw.view(b, hp // ws, ws, wp // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, c).contiguous()
                    items = b * (hp // ws) * (wp // ws)
                    o = self._attention(hw, w, items, ws, heads)
                    o = ops.gemm(o, w["wo"], bias=w["bo"])
                    o = o.view(b, hp // ws, wp // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, hp, wp, c)[:, :gh, :gw]
                    xt = ops.axpby(xt, o.reshape(b * gh * gw, c).contiguous(), 1.0, 1.0)
                else:
                    assert gh == gw, "global attention blocks need a square token grid"
                    o = self._attention(h, w, b, gh, heads)
                    xt = ops.gemm(o, w["wo"], bias=w["bo"], resid=xt)
                h = ops.layernorm(xt, w["g2"], w["be2"], float(blk.norm2.eps))
                h = ops.gemm(h, w["w1"], bias=w["b1"], act=ops.ACT_GELU)
                xt = ops.gemm(h, w["w2"], bias=w["b2"], resid=xt)
            nk = mod.neck
            wn = cached(nk, "sam_neck", [nk[0].weight, nk[1].weight, nk[1].bias, nk[2].weight, nk[3].weight, nk[3].bias],
                        lambda: dict(w0=_bf(nk[0].weight.reshape(nk[0].weight.shape[0], -1)), g1=_bf(nk[1].weight), b1=_bf(nk[1].bias),
                                     w2=ops.pack_conv3x3(nk[2].weight.detach()), g3=_bf(nk[3].weight), b3=_bf(nk[3].bias)))
            eps = float(nk[1].eps)
            y = ops.layernorm(ops.gemm(xt, wn["w0"]), wn["g1"], wn["b1"], eps)        # 1x1 conv, LayerNorm2d = LN over channels
            oc = y.shape[1]
            y = ops.conv3x3(y.view(b, gh, gw, oc), wn["w2"])
            y = ops.layernorm(y.view(b * gh * gw, oc), wn["g3"], wn["b3"], float(nk[3].eps))
            return y.view(b, gh, gw, oc).permute(0, 3, 1, 2).to(pw.dtype)


def wrap_text_encoder(module):
    """The pipeline's text encoder: native when ``module`` is a CLIPTextModel with 64-wide heads living on a GPU."""
    if isinstance(module, ClipTextNative) or module is None:
        return module
    if ClipTextNative.supports(module):
        return _DeviceSwitch(module, ClipTextNative(module))
    return module


class _DeviceSwitch:
    """Calls the native encoder when the wrapped module's parameters are on a CUDA device and the module itself
    otherwise (CPU plumbing tests, configs[0]); attribute access falls through to the module (``.config``, ``.to``)."""

    def __init__(self, module, native):
        self.__dict__["_module"], self.__dict__["_native"] = module, native

    def __call__(self, *a, **k):
        p = next(self._module.parameters(), None)
        if p is not None and p.is_cuda:
            return self._native(*a, **k)
        return self._module(*a, **k)

    def __getattr__(self, name):
        return getattr(self.__dict__["_module"], name)

    def to(self, *a, **k):
        self._module.to(*a, **k)
        return self
