"""Pin the oracle: every function of oracle/ against outputs of the UNMODIFIED reference modules
(tests/golden/*.pt, produced by tools/make_golden.py in the build container).  fp32 CPU on both sides,
so the tolerance only absorbs summation-order differences."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_util import load, synth_state, synth_tensor  # noqa: E402

from oracle import ddim as OD  # noqa: E402
from oracle import geometry as G  # noqa: E402
from oracle import mvgen as OM  # noqa: E402
from oracle import unet3d as OU  # noqa: E402
from oracle import vae as OV  # noqa: E402
from oracle.nn_ops import P  # noqa: E402

torch.set_grad_enabled(False)

TINY = dict(groups=32, heads=(1, 2, 4, 4), mm_heads=2, num_tokens=16)


def close(a, b, tol=2e-4):
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert err <= tol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("tag", ["same", "widen"])
def test_resnet_block(tag):
    g = load("resnet3d.pt")[tag]
    sd = synth_state(g["shapes"], g["seed"])
    x, temb = synth_tensor(g["x_shape"], g["x_seed"]), synth_tensor((2, 48), g["temb_seed"])
    close(OU.resnet_block(x, temb, P(sd), dict(groups=8, resnet_eps=1e-5)), g["y"])


def test_spatial_transformer():
    g = load("transformer3d.pt")
    sd = synth_state(g["shapes"], g["seed"])
    x, ctx = synth_tensor(g["x_shape"], g["x_seed"]), synth_tensor(g["ctx_shape"], g["ctx_seed"])
    y = OU.spatial_transformer(x, ctx, P(sd), 2, dict(groups=8, num_tokens=4, ip_scale=1.0))
    close(y, g["y"])


def test_temporal_module():
    g = load("motion.pt")
    sd = synth_state(g["shapes"], g["seed"])
    x = synth_tensor(g["x_shape"], g["x_seed"])
    close(OU.temporal_module(x, P(sd), dict(groups=32, mm_heads=2, temporal_pe_max_len=64)), g["y"])


def test_geometry_resample_and_masks():
    g = load("geometry.pt")
    cams = g["cams"]
    e_img, p_img = synth_tensor((3, 2, 8, 16), g["e_seed"]), synth_tensor((3, 2, 4, 4), g["p_seed"])
    close(G.e2p(e_img, cams, (4, 4)), g["e2p_bilinear"], 1e-5)
    close(G.e2p(e_img, cams, (4, 4), mode="nearest"), g["e2p_nearest"], 0)
    eq, mask = G.p2e(p_img, cams, (8, 16))
    close(eq, g["p2e"], 1e-5)
    assert torch.equal(mask, g["p2e_mask"])
    pm, em = G.raw_masks(4, 4, 8, 16, cams, "cpu", torch.float32, antipodal=False)
    close(pm, g["raw_pers"], 1e-5); close(em, g["raw_equi"], 1e-5)
    pm, em = G.raw_masks(4, 4, 8, 16, cams, "cpu", torch.float32, antipodal=True)
    close(pm, g["raw_pers_oppo"], 1e-5); close(em, g["raw_equi_oppo"], 1e-5)
    for tag, anti in (("normal", False), ("oppo", True)):
        pm, em = G.merged_masks(4, 4, 8, 16, cams, "cpu", torch.float32, anti)
        close(pm, g[f"merged_pers_{tag}"], 1e-5); close(em, g[f"merged_equi_{tag}"], 1e-5)
    pc, ec = G.polar_coords(4, 4, 8, 16, cams, "cpu", torch.float32)
    close(pc, g["pers_coords"], 1e-6); close(ec, g["equi_coords"], 1e-6)
    th, ph = G.icosahedron_cameras()
    close(torch.tensor(th), g["ico_theta"], 1e-12); close(torch.tensor(ph), g["ico_phi"], 1e-12)


@pytest.mark.parametrize("tag,anti", [("normal", False), ("oppo", True)])
def test_warp_attn(tag, anti):
    g = load("warpattn.pt")
    sd = synth_state(g["shapes"], g["seed"])
    pers, equi = synth_tensor((6, 64, 2, 4, 4), g["pers_seed"]), synth_tensor((2, 64, 2, 8, 16), g["equi_seed"])
    po, eo = OM.warp_attn(pers, equi, g["cams"], P(sd), anti)
    close(po, g[f"pers_{tag}"]); close(eo, g[f"equi_{tag}"])


def test_adapter():
    g = load("adapter.pt")
    feats = synth_tensor((2, 16, 64, 16), g["feats_seed"])
    cfg = dict(tproj_heads=8, adapter_heads=12, adapter_dim_head=64)
    y1 = OU.temporal_projection(feats, P(synth_state(g["tp_shapes"], g["tp_seed"])), cfg)
    close(y1, g["tproj"])
    b, f, n, d = y1.shape
    close(OU.resampler(y1.reshape(b, f * n, d), P(synth_state(g["rs_shapes"], g["rs_seed"])), cfg), g["tokens"])


def test_ddim():
    g = load("ddim.pt")
    s = OD.DDIM()
    close(s.alphas_cumprod, g["alphas_cumprod"], 1e-6)
    for n in (50, 25):
        assert torch.equal(s.set_timesteps(n), g[f"timesteps_{n}"])
    ts = s.set_timesteps(50)
    x = synth_tensor((1, 4, 2, 4, 8), 71)
    for i, t in enumerate(ts):
        x = s.step(synth_tensor(x.shape, 1000 + i), int(t), x)
        close(x, g["traj"][i], 1e-5)
    xb = synth_tensor((1, 4, 2, 4, 8), 71).bfloat16()
    for i, t in enumerate(ts[:5]):
        xb = s.step(synth_tensor(xb.shape, 1000 + i).bfloat16(), int(t), xb)
    assert xb.dtype == torch.bfloat16 and torch.equal(xb, g["bf16_after5"])


def test_vae():
    g = load("vae.pt")
    sd = synth_state(g["shapes"], g["seed"])
    img, z = synth_tensor((2, 3, 32, 48), g["img_seed"]), synth_tensor((2, 4, 4, 6), g["z_seed"])
    close(OV.encode_moments(sd, img, groups=8), g["moments"])
    close(OV.decode(sd, z, groups=8), g["dec"])


def test_unet3d_forward():
    g = load("unet3d.pt")
    sd = synth_state(g["shapes"], g["seed"])
    x, ctx = synth_tensor((1, 9, 4, 8, 16), g["x_seed"]), synth_tensor((1, 21, 32), g["ctx_seed"])
    y = OU.unet3d_forward(sd, x, torch.tensor([g["t"]]), ctx, cfg=TINY, fps=torch.tensor([g["fps"]]))
    close(y, g["y"], 5e-4)


def mvgen_inputs(g):
    s = g["seeds"]
    b, m, f = 2, 2, 16
    rel = torch.tensor([1.0, 1.0, 63.0, 63.0, 128.0, 256.0])[None, None].repeat(b, f, 1)
    pitch = torch.linspace(-5, 5, f)[None].repeat(b, 1)
    return dict(latents=synth_tensor((b, m, 9, f, 16, 16), s["lat"]), pano_latent=synth_tensor((b, 9, f, 32, 64), s["plat"]),
                prompt_embd=synth_tensor((b * m, 5, 32), s["txt_pers"]), pano_prompt_embd=synth_tensor((b, 5, 32), s["txt_pano"]),
                feats_pano=synth_tensor((b, f, 4096, 8), s["feats_pano"]),
                feats_pers=synth_tensor((b, 1, f, 4096, 8), s["feats_pers"]).repeat(1, m, 1, 1, 1),
                rel_pos=rel, pitch=pitch, ip_noise_pano=synth_tensor((b, 16, 32), s["noise_pano"]),
                ip_noise_pers=synth_tensor((b * m, 16, 32), s["noise_pers"]))


def test_mvgen_forward():
    g = load("mvgen.pt")
    sd = synth_state(g["shapes"], g["seed"])
    i = mvgen_inputs(g)
    ys, yp = OM.mv_forward(sd, i["latents"], i["pano_latent"], torch.tensor([g["t"]]), i["prompt_embd"],
                           i["pano_prompt_embd"], g["cams"], torch.tensor([8, 8]), torch.tensor([[8, 8], [8, 8]]),
                           i["feats_pano"], i["feats_pers"], i["rel_pos"], i["pitch"], g["draws"],
                           i["ip_noise_pano"], i["ip_noise_pers"], cfg=TINY)
    close(ys, g["pers"], 1e-3); close(yp, g["pano"], 1e-3)
