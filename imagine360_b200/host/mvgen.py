"""Host-side mirror of ``src.models.MVGenModel.MultiViewBaseModel`` and ``src.modules.attn_perspano.WarpAttn``.

Same constructor, sub-module names (``unet``, ``pano_unet``, ``cp_blocks_encoder.{0,1,2}``, ``cp_blocks_mid``,
``cp_blocks_decoder.{0,1,2}`` each with ``transformer.*``, ``mv_attn.*``, ``pe.freq_bands``) and 14-kwarg ``forward``
as the reference (MVGenModel.py:16-69); the step runs on the native kernels.
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import forward as Fw
from . import geometry as G
from .unet3d import BF16, GEGLU, _EPOCH, _sig, cached, fused_w, geglu_w, lin_w


# ------------------------------------------------------------------------------------------------------
# parameter containers
# ------------------------------------------------------------------------------------------------------
class _WarpCrossAttention(nn.Module):
    """src/modules/transformer.py:43-74 (to_out is a bare Linear here, zero-initialised)"""

    def __init__(self, dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim)
        nn.init.zeros_(self.to_out.weight)
        nn.init.zeros_(self.to_out.bias)


class _WarpFeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        last = nn.Linear(dim * mult, dim)
        nn.init.zeros_(last.weight)
        nn.init.zeros_(last.bias)
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), last)


class _WarpTransformerBlock(nn.Module):
    """src/modules/transformer.py:130-167"""

    def __init__(self, dim, heads, dim_head):
        super().__init__()
        self.attn1 = _WarpCrossAttention(dim, heads, dim_head)
        self.ff = _WarpFeedForward(dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)


class SphericalPE(nn.Module):
    """src/modules/transformer.py:170-188"""

    def __init__(self, n_freqs):
        super().__init__()
        base = 2 if n_freqs <= 80 else 5000 ** (1 / (n_freqs / 2.5))
        self.register_buffer("freq_bands", base ** torch.linspace(0, n_freqs - 1, n_freqs))


class BiasSlot:
    """One WarpAttn call site of a CUDA-graphed step (host/pipeline.py::StepGraph).  A graph bakes pointers, while the
    reference picks the normal or the antipodal mask per call and per step (random.random() < 0.4, utils.py:15-21): the
    graph reads ``static`` and the chosen variant is copied into it before each replay (only when the choice changed)."""

    def __init__(self):
        self.antipodal = False          # variant requested while the slot is still being recorded (eager warm-up)
        self.variants = {}              # {False / True: (bias_equi_q, bias_pers_q)} -- the geometry module's cached tensors
        self.static = None              # (bias_equi_q, bias_pers_q) buffers the graph reads
        self.current = None

    def freeze(self):
        self.static = tuple(torch.empty_like(t) for t in self.variants[False])

    def select(self, antipodal: bool):
        if self.current != antipodal:
            for dst, src in zip(self.static, self.variants[antipodal]):
                dst.copy_(src)
            self.current = antipodal


class WarpAttn(nn.Module):
    """src/modules/attn_perspano.py:10-99"""

    def __init__(self, dim):
        super().__init__()
        self.transformer = _WarpTransformerBlock(dim, dim // 32, 32)
        self.mv_attn = _WarpTransformerBlock(dim, dim // 32, 32)   # present in checkpoints, never used (:15)
        self.pe = SphericalPE(dim // 4)
        self.dim = dim

    @Fw.traced("WarpAttn")
    def forward_native(self, pers, equi, cameras, clips: int, views: int, frames: int, antipodal: bool):
        """pers [(clips*views*frames), ph, pw, C], equi [(clips*frames), eh, ew, C] (channels-last) -> same shapes."""
        tr = self.transformer
        a = tr.attn1
        c = self.dim
        heads, hd = a.heads, a.dim_head
        _, ph, pw, _ = pers.shape
        _, eh, ew, _ = equi.shape
        hw, en = ph * pw, eh * ew
        if isinstance(antipodal, BiasSlot):
            slot = antipodal
            if slot.static is None:
                slot.variants[slot.antipodal] = G.warp_biases(ph, pw, eh, ew, cameras, pers.device, slot.antipodal)
                bias_e, bias_p = slot.variants[slot.antipodal]
            else:
                bias_e, bias_p = slot.static
        else:
            bias_e, bias_p = G.warp_biases(ph, pw, eh, ew, cameras, pers.device, antipodal)
        pe_p, pe_e = G.spherical_pe_tables(self.pe.freq_bands, ph, pw, eh, ew, cameras, pers.device)
        pt, et = Fw.tokens(pers), Fw.tokens(equi)
        g1, b1 = tr.norm1.weight, tr.norm1.bias
        # norm1(x + pe) serves as the query source of one direction and the key/value source of the other
        pn = ops.layernorm(pt, g1, b1, tr.norm1.eps, pre_add=pe_p, pre_index=(frames * hw, views, hw, hw))
        en_ = ops.layernorm(et, g1, b1, tr.norm1.eps, pre_add=pe_e, pre_index=(1, 1, 0, en))
        wq = lin_w(a.to_q)[0]
        wkv = fused_w(a, "kv", [a.to_k, a.to_v])
        wo, bo = lin_w(a.to_out)
        bf = clips * frames
        # perspective -> equirect
        q_e, kv_p = ops.gemm(en_, wq), ops.gemm(pn, wkv)
        o_e = torch.empty_like(et)
        ops.attention(ops.seq_view(q_e, bf, en), ops.multiview_view(kv_p, clips, views, frames, hw, 0),
                      ops.multiview_view(kv_p, clips, views, frames, hw, c), ops.seq_view(o_e, bf, en), heads, hd, bf, bias=bias_e)
        e_out, e_st = Fw.token_linear(o_e, wo, bo, et, want_stats=Fw.fold_pays(c, ops.ACT_GEGLU))
        # equirect -> perspective
        q_p, kv_e = ops.gemm(pn, wq), ops.gemm(en_, wkv)
        o_p = torch.empty_like(pt)
        ops.attention(ops.multiview_view(q_p, clips, views, frames, hw), ops.seq_view(kv_e, bf, en, 0), ops.seq_view(kv_e, bf, en, c),
                      ops.multiview_view(o_p, clips, views, frames, hw), heads, hd, bf, bias=bias_p)
        p_out, p_st = Fw.token_linear(o_p, wo, bo, pt, want_stats=Fw.fold_pays(c, ops.ACT_GEGLU))
        e_out = _warp_ff(e_out, e_st, tr)
        p_out = _warp_ff(p_out, p_st, tr)
        return p_out.view_as(pers), e_out.view_as(equi)


def _warp_ff(t, st, tr):
    g = Fw.ln_linear(t, st, tr.norm2, tr.ff, "geglu", [tr.ff.net[0].proj], bias_mod=tr.ff.net[0].proj, act=ops.ACT_GEGLU)
    w2, b2 = lin_w(tr.ff.net[2])
    return ops.gemm(g, w2, bias=b2, resid=t)


# ------------------------------------------------------------------------------------------------------
# adapter: TemporalProjection + Resampler (resampler.py) -- step-invariant, computed once per clip
# ------------------------------------------------------------------------------------------------------
def _plain_ff(x, ff, resid):
    ln, l1, l3 = ff[0], ff[1], ff[3]
    h = ops.layernorm(x, ln.weight, ln.bias, ln.eps)
    h = ops.gemm(h, lin_w(l1)[0], act=ops.ACT_GELU)
    return ops.gemm(h, lin_w(l3)[0], resid=resid)


def temporal_projection(tp, feats):
    """TemporalProjection.forward (resampler.py:231-267). feats [b, f, hw, c] -> [b, f/16, hw/16, 4c] bf16."""
    b, f, d, c = feats.shape
    hs = int(d ** 0.5)
    # 4x4/stride-4 patch embedding == GEMM over space-to-depth rows (c, kh, kw) to match the conv weight layout
    x = feats.to(BF16).reshape(b * f, hs // 4, 4, hs // 4, 4, c).permute(0, 1, 3, 5, 2, 4).reshape(b * f * (d // 16), c * 16)
    w = cached(tp.patch_embed, "w", [tp.patch_embed.weight, tp.patch_embed.bias],
               lambda: (tp.patch_embed.weight.to(BF16).reshape(tp.patch_embed.weight.shape[0], -1).contiguous(),
                        tp.patch_embed.bias.to(BF16)))
    t = ops.gemm(x.contiguous(), w[0], bias=w[1])                     # rows (b, f, d')
    d, c = d // 16, t.shape[1]

    def stage(t, f, attn, norm, ff, ffnorm):
        n = ops.layernorm(t, norm.weight, norm.bias, norm.eps)
        inner = attn.heads * attn.dim_head
        qkv = ops.gemm(n, fused_w(attn, "qkv", [attn.to_q, attn.to_k, attn.to_v]))
        o = torch.empty((t.shape[0], inner), dtype=BF16, device=t.device)
        ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], o, b, f, d, attn.heads, attn.dim_head)
        wo, bo = lin_w(attn.to_out[0])
        t = ops.gemm(o, wo, bias=bo, resid=t)
        t = _plain_ff(ops.layernorm(t, ffnorm.weight, ffnorm.bias, ffnorm.eps), ff, t)
        return ops.avgpool_frames4(t.view(b, f, d, c)).view(-1, c), f // 4

    t, f = stage(t, f, tp.attn_temp, tp.norm_temp, tp.ff, tp.norm1)
    t, f = stage(t, f, tp.attn_temp_2, tp.norm_temp_2, tp.ff_2, tp.norm2)
    return t.view(b, f, d, c)


def resampler(rs, x):
    """Resampler.forward (resampler.py:132-160).  x [b, n, emb] -> [b, num_queries, out]."""
    b, n, _ = x.shape
    nq, dim = rs.latents.shape[1], rs.latents.shape[2]
    wi, bi = lin_w(rs.proj_in)
    xt = ops.gemm(x.reshape(b * n, -1), wi, bias=bi)
    lat = rs.latents.to(BF16).repeat(b, 1, 1).reshape(b * nq, dim).contiguous()
    for attn, ff in rs.layers:
        heads, dh = attn.heads, attn.dim_head
        inner = heads * dh
        xn = ops.layernorm(xt, attn.norm1.weight, attn.norm1.bias, attn.norm1.eps)
        ln = ops.layernorm(lat, attn.norm2.weight, attn.norm2.bias, attn.norm2.eps)
        q = ops.gemm(ln, lin_w(attn.to_q)[0])
        kv_in = torch.cat([xn.view(b, n, dim), ln.view(b, nq, dim)], dim=1).reshape(b * (n + nq), dim)
        kv = ops.gemm(kv_in, lin_w(attn.to_kv)[0])
        o = torch.empty((b * nq, inner), dtype=BF16, device=x.device)
        ops.attention(ops.seq_view(q, b, nq), ops.seq_view(kv, b, n + nq, 0), ops.seq_view(kv, b, n + nq, inner),
                      ops.seq_view(o, b, nq), heads, dh, b)
        lat = ops.gemm(o, lin_w(attn.to_out)[0], resid=lat)
        lat = _plain_ff(lat, ff, lat)
    wo, bo = lin_w(rs.proj_out)
    out = ops.gemm(lat, wo, bias=bo)
    return ops.layernorm(out, rs.norm_out.weight, rs.norm_out.bias, rs.norm_out.eps).view(b, nq, -1)


def ip_tokens_clean(unet, feats):
    """temporal_proj -> reshape -> image_proj_model (MVGenModel.py:162-184), noise-free."""
    x = temporal_projection(unet.temporal_proj, feats)
    b, f, n, d = x.shape
    return resampler(unet.image_proj_model, x.reshape(b, f * n, d))


def relpos_tokens(unet, rel_pos, pitch, n_tokens):
    """MVGenModel.py:189-222: per frame [cond_rp_proj(add_cond_embedding(sincos(rel_pos))) | add_cond_embedding2(sincos(pitch))],
    the last frame's row repeated up to n_tokens.  All frames batched into three small GEMMs."""
    b, f, _ = rel_pos.shape
    e1 = unet.add_cond_proj(rel_pos.reshape(-1).float()).reshape(b * f, -1).to(BF16)
    e1 = ops.gemm(Fw._mlp(e1, unet.add_cond_embedding), lin_w(unet.cond_rp_proj)[0])
    e2 = Fw._mlp(unet.add_cond_proj(pitch.reshape(-1).float()).to(BF16), unet.add_cond_embedding2)
    tok = torch.cat([e1, e2], dim=-1).view(b, f, -1)
    if n_tokens > f:
        tok = torch.cat([tok, tok[:, -1:].expand(-1, n_tokens - f, -1)], dim=1)
    return tok.contiguous()


# ------------------------------------------------------------------------------------------------------
# MultiViewBaseModel
# ------------------------------------------------------------------------------------------------------
class MultiViewBaseModel(nn.Module):
    def __init__(self, unet, pano_unet, pano_pad=True, device="cuda"):
        super().__init__()
        self.unet, self.pano_unet, self.pano_pad = unet, pano_unet, pano_pad
        if not pano_pad:
            raise NotImplementedError("pano_pad=False is never used by the reference pipeline")
        self.cp_blocks_encoder = nn.ModuleList([WarpAttn(blk.downsamplers[-1].out_channels)
                                                for blk in unet.down_blocks if blk.downsamplers is not None])
        self.cp_blocks_mid = WarpAttn(unet.mid_block.resnets[-1].out_channels)
        self.cp_blocks_decoder = nn.ModuleList([WarpAttn(blk.upsamplers[0].channels)
                                                for blk in unet.up_blocks if blk.upsamplers is not None])
        self._adapter_cache = {}

    # -- step-invariant adapter tokens ------------------------------------------------------------------
    def _adapter_params(self):
        ps = []
        for u in (self.unet, self.pano_unet):
            for name in ("temporal_proj", "image_proj_model", "add_cond_embedding", "add_cond_embedding2", "cond_rp_proj"):
                mod = getattr(u, name, None)
                if mod is not None:
                    ps.extend(mod.parameters())
        return ps

    def _adapter(self, feats_pano, feats_pers, rel_pos, pitch):
        """Clean IP tokens of both branches, recomputed only when the conditioning tensors or the adapter weights change.
        The cache entry holds strong references to the keyed tensors: while it lives their storage cannot be freed
        and handed to the next clip's tensors, so (data_ptr, _version, shape, strides) identifies the CONTENTS
        (a freed-and-reallocated block with new contents would otherwise reproduce the same key)."""
        keyed = (feats_pano, feats_pers, rel_pos, pitch)
        key = (tuple((t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.dtype) for t in keyed),
               _sig(self._adapter_params()), _EPOCH[0])
        hit = self._adapter_cache.get("k")
        if hit is not None and hit[0] == key:
            return hit[1]
        val = self._adapter_compute(feats_pano, feats_pers, rel_pos, pitch)
        self._adapter_cache["k"] = (key, val, keyed)      # keyed: keeps the storages alive (see above)
        return val

    @Fw.traced("adapter")
    def _adapter_compute(self, feats_pano, feats_pers, rel_pos, pitch):
        b, m = feats_pers.shape[:2]
        ip_pano = ip_tokens_clean(self.pano_unet, feats_pano)
        # the reference feeds the SAME features to every view (pipeline...dual.py:716-717): run distinct clips once
        if feats_pers.stride(1) == 0 or (m > 1 and torch.equal(feats_pers[:, 0], feats_pers[:, -1])
                                         and all(torch.equal(feats_pers[:, 0], feats_pers[:, v]) for v in range(1, m))):
            ip_pers = ip_tokens_clean(self.unet, feats_pers[:, 0].contiguous()).repeat_interleave(m, dim=0)
        else:
            ip_pers = ip_tokens_clean(self.unet, feats_pers.reshape(b * m, *feats_pers.shape[2:]))
        rp = relpos_tokens(self.pano_unet, rel_pos, pitch, ip_pano.shape[1]) if self.pano_unet.use_relative_postions == "WithAdapter" else None
        return ip_pano.contiguous(), ip_pers.contiguous(), rp

    @torch.no_grad()
    def forward(self, latents, pano_latent, timestep, prompt_embd, pano_prompt_embd, cameras, use_fps_condition,
                use_ip_plus_cross_attention, fps_tensor_pano, fps_tensor_pers, reference_images_clip_feat_pano,
                reference_images_clip_feat_pers, relative_position_tensor, pitchs_tensor, antipodal_draws=None,
                ip_noise=None):
        """Reference signature (MVGenModel.py:59-69); runs :meth:`_forward` with ``latents.device`` current."""
        with ops.on_device(latents.device):
            return self._forward(latents, pano_latent, timestep, prompt_embd, pano_prompt_embd, cameras, use_fps_condition,
                                 use_ip_plus_cross_attention, fps_tensor_pano, fps_tensor_pers, reference_images_clip_feat_pano,
                                 reference_images_clip_feat_pers, relative_position_tensor, pitchs_tensor, antipodal_draws,
                                 ip_noise)

    def _forward(self, latents, pano_latent, timestep, prompt_embd, pano_prompt_embd, cameras, use_fps_condition,
                 use_ip_plus_cross_attention, fps_tensor_pano, fps_tensor_pers, reference_images_clip_feat_pano,
                 reference_images_clip_feat_pers, relative_position_tensor, pitchs_tensor, antipodal_draws=None,
                 ip_noise=None, adapter_tokens=None):
        """One dual-branch denoise step (MVGenModel.py:59-481).  ``antipodal_draws`` / ``ip_noise`` let tests inject the
        outcomes of the reference's ``random.random() < 0.4`` (src/utils/utils.py:15) and ``torch.randn_like``
        (MVGenModel.py:12) draws; by default they are drawn here in the reference's order."""
        if not use_ip_plus_cross_attention or not use_fps_condition:
            raise NotImplementedError("the Imagine360 path always runs with the IP adapter and the fps condition")
        pu, qu = self.unet, self.pano_unet
        b, m, _, frames, _, _ = latents.shape
        dev = latents.device
        cams = {k: v.reshape(-1, *v.shape[2:]) if isinstance(v, torch.Tensor) and v.dim() >= 2 else v for k, v in cameras.items()}
        cams = {k: (v[:m] if isinstance(v, torch.Tensor) else v) for k, v in cams.items()}
        # 1.1 time + fps embeddings
        t = timestep.to(dev)[:, None].repeat(b, m)
        silu_p = Fw.time_embedding(pu, t.reshape(-1), fps_tensor_pers.to(dev).reshape(-1).to(BF16))
        silu_q = Fw.time_embedding(qu, t[:, 0], fps_tensor_pano.to(dev).to(BF16).expand(b))
        temb_p, temb_q = Fw.all_temb_projections(pu, silu_p), Fw.all_temb_projections(qu, silu_q)
        x = Fw.conv_in(pu, latents.reshape(b * m, *latents.shape[2:]), False)
        y = Fw.conv_in(qu, pano_latent, True)
        # 1.2 IP tokens: clean tokens cached, fresh noise every step (pano first: MVGenModel.py:186-187)
        if adapter_tokens is not None:           # CUDA-graphed step: the clean tokens come from the adapter graph's buffers
            ip_pano, ip_pers, rp = adapter_tokens
        else:
            ip_pano, ip_pers, rp = self._adapter(reference_images_clip_feat_pano, reference_images_clip_feat_pers,
                                                 relative_position_tensor, pitchs_tensor)
        if ip_noise is None:
            ip_noise = (torch.randn_like(ip_pano), torch.randn_like(ip_pers))
        ip_pano = ops.axpby(ip_pano, ip_noise[0].to(BF16).contiguous(), 1.0, 0.1)
        ip_pers = ops.axpby(ip_pers, ip_noise[1].to(BF16).contiguous(), 1.0, 0.1)
        if rp is not None:
            ip_pano = ops.axpby(ip_pano, rp, 1.0, 1.0)
        ctx_q = Fw.Context(pano_prompt_embd.to(BF16), ip_pano)
        ctx_p = Fw.Context(prompt_embd.to(BF16), ip_pers)
        if antipodal_draws is None:
            antipodal_draws = [random.random() < 0.4 for _ in range(7)]
        draws = list(antipodal_draws)
        g = pu.groups

        def warp(block, x, y):
            return block.forward_native(x, y, cams, b, m, frames, draws.pop(0))

        xs, ys = [x], [y]
        for i, (bp, bq) in enumerate(zip(pu.down_blocks, qu.down_blocks)):
            for j in range(len(bq.resnets)):
                ca = bq.has_cross_attention
                x = Fw.resnet_block(x, bp.resnets[j], temb_p[bp.resnets[j]], frames, g, out_stats=ca)
                y = Fw.resnet_block(y, bq.resnets[j], temb_q[bq.resnets[j]], frames, g, halo=2, out_stats=ca)
                if bq.has_cross_attention:
                    x = Fw.spatial_transformer(x, bp.attentions[j], ctx_p, frames)
                    x = Fw.temporal_module(x, bp.motion_modules[j], frames)
                    y = Fw.spatial_transformer(y, bq.attentions[j], ctx_q, frames)
                    y = Fw.temporal_module(y, bq.motion_modules[j], frames)
                # DownBlock3D's motion modules are skipped by the dual forward (MVGenModel.py:292-303)
                xs.append(x)
                ys.append(y)
            if bq.downsamplers is not None:
                x = Fw.downsample(x, bp.downsamplers[0], False)
                y = Fw.downsample(y, bq.downsamplers[0], True)
                xs.append(x)
                ys.append(y)
                x, y = warp(self.cp_blocks_encoder[i], x, y)
        mp, mq = pu.mid_block, qu.mid_block
        x = Fw.resnet_block(x, mp.resnets[0], temb_p[mp.resnets[0]], frames, g, out_stats=True)
        y = Fw.resnet_block(y, mq.resnets[0], temb_q[mq.resnets[0]], frames, g, halo=2, out_stats=True)
        for i in range(len(mq.attentions)):
            x = Fw.spatial_transformer(x, mp.attentions[i], ctx_p, frames)
            x = Fw.temporal_module(x, mp.motion_modules[i], frames)
            x = Fw.resnet_block(x, mp.resnets[i + 1], temb_p[mp.resnets[i + 1]], frames, g)
            y = Fw.spatial_transformer(y, mq.attentions[i], ctx_q, frames)
            y = Fw.temporal_module(y, mq.motion_modules[i], frames)
            y = Fw.resnet_block(y, mq.resnets[i + 1], temb_q[mq.resnets[i + 1]], frames, g, halo=2)
        x, y = warp(self.cp_blocks_mid, x, y)
        dec = 0
        for bp, bq in zip(pu.up_blocks, qu.up_blocks):
            for j in range(len(bq.resnets)):
                ca = bq.has_cross_attention
                x = Fw.resnet_block(x, bp.resnets[j], temb_p[bp.resnets[j]], frames, g, skip=xs.pop(), out_stats=ca)
                y = Fw.resnet_block(y, bq.resnets[j], temb_q[bq.resnets[j]], frames, g, skip=ys.pop(), halo=2, out_stats=ca)
                if bq.has_cross_attention:
                    x = Fw.spatial_transformer(x, bp.attentions[j], ctx_p, frames)
                    x = Fw.temporal_module(x, bp.motion_modules[j], frames)
                    y = Fw.spatial_transformer(y, bq.attentions[j], ctx_q, frames)
                    y = Fw.temporal_module(y, bq.motion_modules[j], frames)
            if bq.upsamplers is not None:
                x, y = warp(self.cp_blocks_decoder[dec], x, y)
                dec += 1
                x = Fw.upsample(x, bp.upsamplers[0], False)
                y = Fw.upsample(y, bq.upsamplers[0], True)
        sample = Fw.conv_out(pu, x, b * m, False)
        sample = sample.reshape(b, m, *sample.shape[1:])
        pano_sample = Fw.conv_out(qu, y, b, True)
        return sample, pano_sample
