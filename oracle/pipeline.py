"""Oracle: the DDIM loop of AnimationPipeline.__call__ and its noise/latent preparation.

Follows animatediff/pipelines/pipeline_animation_inference_dual.py:361-387 (init_noise), :734-809 (loop),
:301-313/:811-815 (circular-padded VAE decode) of the reference."""
from __future__ import annotations

import random

import torch

from . import geometry as G
from . import vae as V
from .ddim import DDIM, cfg_combine
from .mvgen import mv_forward


def init_noise(pano_noise, cameras, pers_hw, dtype):
    """init_noise (:361-387).  pano_noise: the randn draw [b, f, 1, 4, eh, ew] (fp32).  Every view receives the
    panorama noise resampled with NEAREST lookup, so the two branches start from correlated noise."""
    b, f = pano_noise.shape[:2]
    m = len(cameras["FoV"])
    pano_out = pano_noise.squeeze(2).permute(0, 2, 1, 3, 4)           # b c f h w
    views = []
    for i in range(f):
        frame = pano_noise[:, i].expand(-1, m, -1, -1, -1).reshape(b * m, *pano_noise.shape[3:])
        cams = {k: list(v) * b for k, v in cameras.items()}
        v = G.e2p(frame, cams, pers_hw, mode="nearest")
        views.append(v.reshape(b, m, *v.shape[1:]))
    pers = torch.stack(views, dim=0).permute(1, 2, 3, 0, 4, 5)       # f b m c h w -> b m c f h w
    return pano_out.to(dtype), pers.to(dtype)


def denoise_loop(sd, pano_latent, pers_latent, pano_mask, pers_masks, pano_masked, pers_masked, cond, cameras,
                 num_steps, guidance=7.5, cfg=None, seed_python=None, mask_dtype=None, noise_fn=None, grid_dtype=None,
                 pe_dtype=None):
    """The `for t in timesteps` loop (:734-809).  ``cond`` holds the step-invariant conditioning:
    text_pano [2,77,D], text_pers [2m,77,D], feats_pano [2,f,4096,C], feats_pers [2,m,f,4096,C], fps (int),
    rel_pos [f,6], pitch [f].  Python's ``random`` supplies the 7 antipodal draws per step in the order
    enc0, enc1, enc2, mid, dec0, dec1, dec2; torch's global RNG supplies the two IP-token noise draws per step
    (pano first), unless ``noise_fn(shape)`` is given.  ``grid_dtype`` / ``pe_dtype``: see mvgen.warp_attn."""
    sched = DDIM()
    ts = sched.set_timesteps(num_steps)
    if seed_python is not None:
        random.seed(seed_python)
    m = pers_latent.shape[1]
    dtype = pano_latent.dtype
    fps_pano = torch.tensor([cond["fps"]] * 2, device=pano_latent.device)
    fps_pers = fps_pano[:, None].repeat(1, m)
    rel_pos = cond["rel_pos"][None].repeat(2, 1, 1)
    pitch = cond["pitch"][None].repeat(2, 1)
    ntok = cfg.get("num_tokens", 64) if cfg else 64
    dctx = cond["text_pano"].shape[-1]
    randn = noise_fn or (lambda shape: torch.randn(shape, dtype=dtype, device=pano_latent.device))
    for t in ts:
        xin_pano = torch.cat([pano_latent, pano_mask, pano_masked], dim=1)
        xin_pers = torch.cat([pers_latent, pers_masks, pers_masked], dim=2)
        draws = [random.random() < 0.4 for _ in range(7)]
        n_pano = randn((2, ntok, dctx))
        n_pers = randn((2 * m, ntok, dctx))
        pred_pers, pred_pano = mv_forward(
            sd, torch.cat([xin_pers] * 2), torch.cat([xin_pano] * 2), t.reshape(1).to(pano_latent.device), cond["text_pers"], cond["text_pano"],
            cameras, fps_pano, fps_pers, cond["feats_pano"], cond["feats_pers"], rel_pos, pitch, draws, n_pano, n_pers,
            cfg=cfg, mask_dtype=mask_dtype, grid_dtype=grid_dtype, pe_dtype=pe_dtype)
        pano_latent = sched.step(cfg_combine(pred_pano, guidance), int(t), pano_latent)
        pers_latent = sched.step(cfg_combine(pred_pers, guidance), int(t), pers_latent)
    return pano_latent, pers_latent


def decode_video(vae_sd, pano_latent, groups=32):
    """pad_pano(latent, 4) -> per-frame vae.decode -> /2+0.5 clamp -> crop 32 px (:811-815, :301-313).
    Returns fp32 [b, 3, f, H, W] in [0, 1]."""
    z = G.pad_pano(pano_latent, 4) / 0.18215
    b, c, f, h, w = z.shape
    frames = z.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    out = torch.cat([V.decode(vae_sd, frames[i:i + 1], groups) for i in range(b * f)])
    out = out.reshape(b, f, 3, out.shape[-2], out.shape[-1]).permute(0, 2, 1, 3, 4)
    out = (out / 2 + 0.5).clamp(0, 1).float()
    return G.unpad_pano(out, 32)
