"""Oracle: WarpAttn and the dual-branch MultiViewBaseModel.forward.

Follows src/modules/attn_perspano.py:22-99, src/modules/transformer.py:43-167 and
src/models/MVGenModel.py:59-481 of the reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import geometry as G
from .nn_ops import P, attention_core, conv2d_frames, group_norm_frames, linear, timestep_embedding, timestep_mlp
from .unet3d import (DEFAULT_CFG, downsample, resampler, resnet_block, spatial_transformer, temporal_module,
                     temporal_projection, time_embedding, upsample)


# --------------------------------------------------------------------------------------------
# WarpAttn
# --------------------------------------------------------------------------------------------
def warp_transformer(x, context, mask, query_pe, p: P):
    """BasicTransformerBlock._forward (src/modules/transformer.py:151-167): q and kv share norm1, the
    residual adds the un-PE'd query; CrossAttention (:59-74) uses only mask[0], broadcast over
    batch x heads; head_dim 32."""
    c = x.shape[-1]
    heads = c // 32
    ap = p.sub("attn1")
    q_in = F.layer_norm(x + query_pe, (c,), p["norm1.weight"], p["norm1.bias"])
    kv_in = F.layer_norm(context, (c,), p["norm1.weight"], p["norm1.bias"])
    q, k, v = linear(q_in, ap, "to_q"), linear(kv_in, ap, "to_k"), linear(kv_in, ap, "to_v")
    o = attention_core(q, k, v, heads, bias=mask[0].to(q.dtype))
    x = linear(o, ap, "to_out") + x
    fp = p.sub("ff")
    h = linear(F.layer_norm(x, (c,), p["norm2.weight"], p["norm2.bias"]), fp, "net.0.proj")
    val, gate = h.chunk(2, dim=-1)
    return linear(val * F.gelu(gate), fp, "net.2") + x


def warp_attn(pers_x, equi_x, cameras, p: P, antipodal: bool, mask_dtype=None, grid_dtype=None, pe_dtype=None):
    """WarpAttn.forward (attn_perspano.py:22-99).  pers_x [(b m), c, f, ph, pw], equi_x [b, c, f, eh, ew].
    ``antipodal`` is the outcome of the reference's ``random.random() < 0.4`` draw (utils.py:15-21).
    ``grid_dtype`` / ``pe_dtype``: dtype in which the sampling grids and the spherical PE are evaluated; the reference
    uses the activation dtype for both (bf16 in production), which matters: the PE frequencies reach 2^79 and the
    grids hold pixel coordinates, so an fp32 evaluation of the rest must still quantise these two like production."""
    bm, c, f, ph, pw = pers_x.shape
    b, _, _, eh, ew = equi_x.shape
    m = bm // b
    dt = pers_x.dtype
    mdt = mask_dtype or dt
    pers_masks, equi_masks = G.merged_masks(ph, pw, eh, ew, cameras, pers_x.device, mdt, antipodal, grid_dtype)
    if grid_dtype is not None:      # production rounds the finished bias to the activation dtype
        pers_masks, equi_masks = pers_masks.to(grid_dtype).to(mdt), equi_masks.to(grid_dtype).to(mdt)
    pdt = pe_dtype or dt
    pers_coords, equi_coords = G.polar_coords(ph, pw, eh, ew, cameras, pers_x.device, pdt)
    fb = p.get("pe.freq_bands")
    if fb is None:
        fb = G.spherical_freqs(c // 4, pers_x.device)
    pers_pe = G.spherical_pe(pers_coords, fb.to(pdt)).to(dt)   # [m, ph, pw, c]
    equi_pe = G.spherical_pe(equi_coords, fb.to(pdt)).to(dt)   # [eh, ew, c]

    # tokens: equi '(b f) (h w) c', pers '(b f) (m h w) c'
    def equi_tokens(t):
        return t.permute(0, 2, 3, 4, 1).reshape(b * f, eh * ew, c)

    def pers_tokens(t):
        return t.reshape(b, m, c, f, ph, pw).permute(0, 3, 1, 4, 5, 2).reshape(b * f, m * ph * pw, c)

    equi_pe_tok = equi_pe.reshape(1, eh * ew, c).expand(b * f, -1, -1)
    pers_pe_tok = pers_pe.reshape(1, m * ph * pw, c).expand(b * f, -1, -1)
    equi_tok, pers_tok = equi_tokens(equi_x), pers_tokens(pers_x)
    tp = p.sub("transformer")
    # perspective -> equirect: q = equi, kv = pers (+PE), bias[(eh ew), (m ph pw)]
    bias_e = pers_masks.permute(1, 2, 0, 3, 4).reshape(1, eh * ew, m * ph * pw)
    equi_out = warp_transformer(equi_tok, pers_tok + pers_pe_tok, bias_e, equi_pe_tok, tp)
    # equirect -> perspective: q = pers, kv = equi (+PE), bias[(m ph pw), (eh ew)]
    bias_p = equi_masks.reshape(1, m * ph * pw, eh * ew)
    pers_out = warp_transformer(pers_tok, equi_tok + equi_pe_tok, bias_p, pers_pe_tok, tp)
    pers_out = pers_out.reshape(b, f, m, ph, pw, c).permute(0, 2, 5, 1, 3, 4).reshape(bm, c, f, ph, pw)
    equi_out = equi_out.reshape(b, f, eh, ew, c).permute(0, 4, 1, 2, 3)
    return pers_out, equi_out


# --------------------------------------------------------------------------------------------
# adapter tokens (MVGenModel.py:155-246)
# --------------------------------------------------------------------------------------------
def ip_tokens_clean(feats, p: P, cfg):
    """temporal_proj -> reshape -> image_proj_model (MVGenModel.py:162-184). feats [b, f, hw, c]."""
    x = temporal_projection(feats, p.sub("temporal_proj"), cfg)
    b, f, n, d = x.shape
    return resampler(x.reshape(b, f * n, d), p.sub("image_proj_model"), cfg)


def relpos_tokens(rel_pos, pitch, p: P, dtype, time_dim, n_tokens):
    """MVGenModel.py:189-222 (pano branch only): per-frame [cond_rp_proj(add_cond_embedding(sincos(rel_pos[6]))) |
    add_cond_embedding2(sincos(pitch))], the last frame's vector repeated up to n_tokens."""
    b, f = rel_pos.shape[:2]
    out = []
    for i in range(f):
        e1 = timestep_embedding(rel_pos[:, i, :].flatten(), time_dim).reshape(b, -1).to(dtype)
        e1 = F.linear(timestep_mlp(e1, p.sub("add_cond_embedding")), p["cond_rp_proj.weight"])
        e2 = timestep_embedding(pitch[:, i].flatten(), time_dim).reshape(b, -1).to(dtype)
        e2 = timestep_mlp(e2, p.sub("add_cond_embedding2"))
        out.append(torch.cat([e1, e2], dim=-1))
    out += [out[-1]] * (n_tokens - f)
    return torch.stack(out, dim=1)


# --------------------------------------------------------------------------------------------
# MultiViewBaseModel.forward
# --------------------------------------------------------------------------------------------
def _pano_resnet(x, emb, p, cfg):
    """pad_pano(2) -> ResnetBlock3D -> unpad_pano(2) (MVGenModel.py:276-281): GroupNorm statistics are
    those of the padded tensor (SURVEY.md trap 1)."""
    return G.unpad_pano(resnet_block(G.pad_pano(x, 2), emb, p, cfg), 2)


def mv_forward(sd, latents, pano_latent, timestep, prompt_embd, pano_prompt_embd, cameras, fps_pano, fps_pers,
               feats_pano, feats_pers, rel_pos, pitch, antipodal_draws, ip_noise_pano, ip_noise_pers, cfg=None,
               mask_dtype=None, grid_dtype=None, pe_dtype=None):
    """One dual-branch denoise step.

    latents [b, m, 9, f, h, w]; pano_latent [b, 9, f, H, W]; timestep [1] int64; prompt_embd [b*m, 77, D];
    pano_prompt_embd [b, 77, D]; feats_* SAM features [b, (m,) f, 4096, Ch]; rel_pos [b, f, 6]; pitch [b, f].
    ``antipodal_draws``: 7 booleans (enc0, enc1, enc2, mid, dec0, dec1, dec2) -- the outcomes of the
    reference's random.random() < 0.4.  ``ip_noise_*``: the randn_like draws of add_noise_to_condition
    (MVGenModel.py:11-14,:186-187), already multiplied by nothing (this function applies the 0.1)."""
    cfg = {**DEFAULT_CFG, **(cfg or {})}
    pp, qp = P(sd, "unet."), P(sd, "pano_unet.")
    dtype = pano_latent.dtype
    b, m, c9, f, h, w = latents.shape
    x = latents.reshape(b * m, c9, f, h, w)
    time_dim = pp["conv_in.weight"].shape[0]
    t_pers = timestep[:, None].repeat(b, m).reshape(-1)
    t_pano = timestep[:, None].repeat(b, m)[:, 0]
    emb = time_embedding(t_pers, fps_pers.reshape(-1).to(dtype), pp, dtype, time_dim)
    pano_emb = time_embedding(t_pano, fps_pano.to(dtype).expand(b), qp, dtype, time_dim)

    x = conv2d_frames(x, pp, "conv_in")
    y = G.unpad_pano(conv2d_frames(G.pad_pano(pano_latent, 1), qp, "conv_in"), 1)

    # adapter
    ip_pano = ip_tokens_clean(feats_pano, qp, cfg)
    ip_pers = ip_tokens_clean(feats_pers.reshape(b * m, *feats_pers.shape[2:]), pp, cfg)
    ip_pano = ip_pano + 0.1 * ip_noise_pano
    ip_pers = ip_pers + 0.1 * ip_noise_pers
    ip_pano = ip_pano + relpos_tokens(rel_pos, pitch, qp, dtype, time_dim, ip_pano.shape[1])
    pano_ctx = torch.cat([pano_prompt_embd, ip_pano], dim=1)
    pers_ctx = torch.cat([prompt_embd, ip_pers], dim=1)

    draws = list(antipodal_draws)
    xs, ys = [x], [y]
    heads = cfg["heads"]
    nd = qp.count("down_blocks")
    for i in range(nd):
        bp, bq = pp.sub(f"down_blocks.{i}"), qp.sub(f"down_blocks.{i}")
        has_attn = bq.has_prefix("attentions")
        for j in range(bq.count("resnets")):
            x = resnet_block(x, emb, bp.sub(f"resnets.{j}"), cfg)
            y = _pano_resnet(y, pano_emb, bq.sub(f"resnets.{j}"), cfg)
            if has_attn:
                x = spatial_transformer(x, pers_ctx, bp.sub(f"attentions.{j}"), heads[i], cfg)
                x = temporal_module(x, bp.sub(f"motion_modules.{j}"), cfg)
                y = spatial_transformer(y, pano_ctx, bq.sub(f"attentions.{j}"), heads[i], cfg)
                y = temporal_module(y, bq.sub(f"motion_modules.{j}"), cfg)
            # DownBlock3D: motion modules are NOT called in the dual path (MVGenModel.py:292-303)
            xs.append(x)
            ys.append(y)
        if bq.has_prefix("downsamplers"):
            x = downsample(x, bp.sub("downsamplers.0"))
            y = G.unpad_pano(downsample(G.pad_pano(y, 2), bq.sub("downsamplers.0")), 1)
            xs.append(x)
            ys.append(y)
            x, y = warp_attn(x, y, cameras, P(sd, f"cp_blocks_encoder.{i}."), draws.pop(0), mask_dtype, grid_dtype, pe_dtype)

    mp, mq = pp.sub("mid_block"), qp.sub("mid_block")
    x = resnet_block(x, emb, mp.sub("resnets.0"), cfg)
    y = _pano_resnet(y, pano_emb, mq.sub("resnets.0"), cfg)
    for i in range(mq.count("attentions")):
        x = spatial_transformer(x, pers_ctx, mp.sub(f"attentions.{i}"), heads[-1], cfg)
        x = temporal_module(x, mp.sub(f"motion_modules.{i}"), cfg)
        x = resnet_block(x, emb, mp.sub(f"resnets.{i + 1}"), cfg)
        y = spatial_transformer(y, pano_ctx, mq.sub(f"attentions.{i}"), heads[-1], cfg)
        y = temporal_module(y, mq.sub(f"motion_modules.{i}"), cfg)
        y = _pano_resnet(y, pano_emb, mq.sub(f"resnets.{i + 1}"), cfg)
    x, y = warp_attn(x, y, cameras, P(sd, "cp_blocks_mid."), draws.pop(0), mask_dtype, grid_dtype, pe_dtype)

    nu = qp.count("up_blocks")
    dec = 0
    for i in range(nu):
        bp, bq = pp.sub(f"up_blocks.{i}"), qp.sub(f"up_blocks.{i}")
        has_attn = bq.has_prefix("attentions")
        for j in range(bq.count("resnets")):
            x = resnet_block(torch.cat([x, xs.pop()], dim=1), emb, bp.sub(f"resnets.{j}"), cfg)
            y = _pano_resnet(torch.cat([y, ys.pop()], dim=1), pano_emb, bq.sub(f"resnets.{j}"), cfg)
            if has_attn:
                x = spatial_transformer(x, pers_ctx, bp.sub(f"attentions.{j}"), heads[nu - 1 - i], cfg)
                x = temporal_module(x, bp.sub(f"motion_modules.{j}"), cfg)
                y = spatial_transformer(y, pano_ctx, bq.sub(f"attentions.{j}"), heads[nu - 1 - i], cfg)
                y = temporal_module(y, bq.sub(f"motion_modules.{j}"), cfg)
        if bq.has_prefix("upsamplers"):
            x, y = warp_attn(x, y, cameras, P(sd, f"cp_blocks_decoder.{dec}."), draws.pop(0), mask_dtype, grid_dtype, pe_dtype)
            dec += 1
            x = upsample(x, bp.sub("upsamplers.0"))
            y = G.unpad_pano(upsample(G.pad_pano(y, 1), bq.sub("upsamplers.0")), 2)

    g, eps = cfg["groups"], cfg["resnet_eps"]
    x = conv2d_frames(F.silu(group_norm_frames(x, pp, "conv_norm_out", g, eps)), pp, "conv_out")
    x = x.reshape(b, m, *x.shape[1:])
    y = F.silu(group_norm_frames(y, qp, "conv_norm_out", g, eps))
    y = G.unpad_pano(conv2d_frames(G.pad_pano(y, 1), qp, "conv_out"), 1)
    return x, y
