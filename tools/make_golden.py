"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported through
tools/ref_shim.py) on CPU in fp32 with deterministic synthetic weights and inputs.

    python tools/make_golden.py [name ...]

/root/reference only exists in the build container, so this script cannot run on the GPU box; the
fixtures it writes are committed.  Fixtures hold parameter SHAPES + seeds + reference outputs only.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_shim  # noqa: E402

ref_shim.install()
from golden_util import GOLDEN_DIR, synth_state, synth_tensor, tiny_cameras  # noqa: E402

torch.set_grad_enabled(False)
os.makedirs(GOLDEN_DIR, exist_ok=True)


def shapes_of(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def load_synth(module, seed):
    shapes = shapes_of(module)
    sd = synth_state(shapes, seed)
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith(("pos_encoder.pe", "pe.freq_bands")) for k in missing), missing
    return shapes


def enable_xformers_like_gpu(module):
    """inference_dual_p2e.py:237 calls enable_xformers_memory_efficient_attention(); on a GPU that flips
    IPCrossAttention._use_memory_efficient_attention_xformers (animatediff/models/attention.py:434-459), whose
    FMHA applies the 1/sqrt(d) softmax scale.  NOTE the quirk this hides: IPCrossAttention.__init__ overwrites
    Attention.scale (= dim_head**-0.5) with the IP-adapter scale 1.0 (attention.py:52), so the *math* path
    (the only one a CPU run would take) silently drops the softmax scale.  The production path is the FMHA one,
    so the fixtures are generated with the flag set, exactly as on the GPU."""
    from animatediff.models.attention import IPCrossAttention
    for m in module.modules():
        if isinstance(m, IPCrossAttention):
            m._use_memory_efficient_attention_xformers = True
    return module


def save(name, obj):
    path = os.path.join(GOLDEN_DIR, name)
    torch.save(obj, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def cam_tensors(cams):
    return {k: torch.tensor(v) for k, v in cams.items()}


# ------------------------------------------------------------------------------------------------
def g_resnet():
    from animatediff.models.resnet import ResnetBlock3D
    out = {}
    for tag, (cin, cout) in {"same": (32, 32), "widen": (32, 64)}.items():
        m = ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=48, groups=8, eps=1e-5,
                          use_inflated_groupnorm=True).eval()
        shapes = load_synth(m, 11)
        x = synth_tensor((2, cin, 3, 6, 10), 12)
        temb = synth_tensor((2, 48), 13)
        out[tag] = dict(shapes=shapes, seed=11, x_seed=12, temb_seed=13, x_shape=list(x.shape), y=m(x, temb))
    save("resnet3d.pt", out)


def g_transformer():
    from animatediff.models.attention import Transformer3DModel
    m = Transformer3DModel(2, 16, in_channels=32, num_layers=1, cross_attention_dim=24, norm_num_groups=8,
                           use_linear_projection=True, upcast_attention=True, unet_use_cross_frame_attention=False,
                           unet_use_temporal_attention=False, use_ip_plus_cross_attention=True,
                           image_cross_attention_dim=24, scale=1.0, num_tokens=4).eval()
    enable_xformers_like_gpu(m)
    shapes = load_synth(m, 21)
    x = synth_tensor((2, 32, 3, 4, 6), 22)
    ctx = synth_tensor((2, 7 + 4, 24), 23)
    save("transformer3d.pt", dict(shapes=shapes, seed=21, x_seed=22, ctx_seed=23, x_shape=list(x.shape),
                                  ctx_shape=list(ctx.shape), y=m(x, encoder_hidden_states=ctx).sample))


def g_motion():
    from animatediff.models.motion_module import VanillaTemporalModule
    m = VanillaTemporalModule(in_channels=32, num_attention_heads=2, num_transformer_block=1,
                              attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=64,
                              temporal_attention_dim_div=1, zero_initialize=True).eval()
    shapes = load_synth(m, 31)
    x = synth_tensor((2, 32, 5, 3, 4), 32)
    save("motion.pt", dict(shapes=shapes, seed=31, x_seed=32, x_shape=list(x.shape), y=m(x, None, None)))


def g_geometry():
    from src.utils.Perspective_and_Equirectangular import e2p, p2e
    from src.utils.utils import get_coords, get_masks, get_oppo_masks
    import src.utils.utils as U
    cams = tiny_cameras(3)
    ct = cam_tensors(cams)
    e_img = synth_tensor((3, 2, 8, 16), 41)
    p_img = synth_tensor((3, 2, 4, 4), 42)
    out = dict(cams=cams, e_seed=41, p_seed=42)
    out["e2p_bilinear"] = e2p(e_img, ct["FoV"], ct["theta"], ct["phi"], (4, 4))
    out["e2p_nearest"] = e2p(e_img, ct["FoV"], ct["theta"], ct["phi"], (4, 4), mode="nearest")
    eq, mask = p2e(p_img, ct["FoV"], ct["theta"], ct["phi"], (8, 16))
    out["p2e"], out["p2e_mask"] = eq, mask
    pm, em = get_masks(4, 4, 8, 16, ct, "cpu", torch.float32)
    out["raw_pers"], out["raw_equi"] = pm, em
    pm, em = get_oppo_masks(4, 4, 8, 16, ct, "cpu", torch.float32)
    out["raw_pers_oppo"], out["raw_equi_oppo"] = pm, em
    for tag, val in (("normal", 0.9), ("oppo", 0.1)):
        U.random.random = lambda v=val: v
        pm, em = U.get_merged_masks(4, 4, 8, 16, ct, "cpu", torch.float32)
        out[f"merged_pers_{tag}"], out[f"merged_equi_{tag}"] = pm, em
    import random as _r
    U.random.random = _r.random
    pc, ec = get_coords(4, 4, 8, 16, ct, "cpu", torch.float32)
    out["pers_coords"], out["equi_coords"] = pc, ec
    # the full 20-camera icosahedron layout used by inference_dual_p2e.get_cameras
    from src.utils.pano import icosahedron_sample_camera
    th, ph = icosahedron_sample_camera()
    out["ico_theta"], out["ico_phi"] = torch.tensor(np.rad2deg(th)), torch.tensor(np.rad2deg(ph))
    save("geometry.pt", out)


def g_warp():
    from src.modules.attn_perspano import WarpAttn
    import src.utils.utils as U
    m = WarpAttn(64).eval()
    shapes = load_synth(m, 51)
    cams = tiny_cameras(3)
    ct = cam_tensors(cams)
    pers = synth_tensor((2 * 3, 64, 2, 4, 4), 52)
    equi = synth_tensor((2, 64, 2, 8, 16), 53)
    out = dict(shapes=shapes, seed=51, cams=cams, pers_seed=52, equi_seed=53)
    for tag, val in (("normal", 0.9), ("oppo", 0.1)):
        U.random.random = lambda v=val: v
        po, eo = m(pers, equi, ct)
        out[f"pers_{tag}"], out[f"equi_{tag}"] = po, eo
    import random as _r
    U.random.random = _r.random
    save("warpattn.pt", out)


def g_adapter():
    from animatediff.models.resampler import Resampler, TemporalProjection
    tp = TemporalProjection(dim=16, dim_head=64, heads=8, compress_video_features=True).eval()
    rs = Resampler(dim=48, depth=4, dim_head=64, heads=12, num_queries=8, embedding_dim=64, output_dim=48,
                   ff_mult=4).eval()
    s1, s2 = load_synth(tp, 61), load_synth(rs, 62)
    feats = synth_tensor((2, 16, 64, 16), 63)
    y1 = tp(feats)
    b, f, n, d = y1.shape
    y2 = rs(y1.reshape(b, f * n, d))
    save("adapter.pt", dict(tp_shapes=s1, rs_shapes=s2, tp_seed=61, rs_seed=62, feats_seed=63, tproj=y1, tokens=y2))


def g_ddim():
    from diffusers import DDIMScheduler
    s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
                      steps_offset=1, clip_sample=False, prediction_type="v_prediction", rescale_betas_zero_snr=True)
    out = dict(alphas_cumprod=s.alphas_cumprod.clone())
    for n in (50, 25):
        s.set_timesteps(n)
        out[f"timesteps_{n}"] = s.timesteps.clone()
    s.set_timesteps(50)
    x = synth_tensor((1, 4, 2, 4, 8), 71)
    traj = []
    for i, t in enumerate(s.timesteps):
        v = synth_tensor(x.shape, 1000 + i)
        x = s.step(v, t, x, eta=0.0).prev_sample
        traj.append(x.clone())
    out["traj"] = torch.stack(traj)
    xb = synth_tensor((1, 4, 2, 4, 8), 71).bfloat16()
    for i, t in enumerate(s.timesteps[:5]):
        xb = s.step(synth_tensor(xb.shape, 1000 + i).bfloat16(), t, xb, eta=0.0).prev_sample
    out["bf16_after5"] = xb
    save("ddim.pt", out)


def g_vae():
    from diffusers import AutoencoderKL
    m = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                      up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(16, 32, 64, 64),
                      layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=8, sample_size=64).eval()
    shapes = load_synth(m, 81)
    img = synth_tensor((2, 3, 32, 48), 82)
    z = synth_tensor((2, 4, 4, 6), 83)
    post = m.encode(img, 2).latent_dist
    save("vae.pt", dict(shapes=shapes, seed=81, img_seed=82, z_seed=83, moments=post.parameters, dec=m.decode(z).sample))


def tiny_unet(seed):
    from animatediff.models.unet import UNet3DConditionModel
    kw = ref_shim.tiny_unet_kwargs(c0=32, heads=(1, 2, 4, 4), cross_dim=32, img_hidden=8, num_tokens=16, mm_heads=2)
    m = enable_xformers_like_gpu(UNet3DConditionModel(**kw).eval())
    shapes = load_synth(m, seed)
    return m, shapes


def g_unet():
    m, shapes = tiny_unet(91)
    x = synth_tensor((1, 9, 4, 8, 16), 92)
    ctx = synth_tensor((1, 5 + 16, 32), 93)
    t = torch.tensor([481])
    y = m(x, t, ctx, use_ip_plus_cross_attention=False, use_fps_condition=True, fps_tensor=torch.tensor([8])).sample
    save("unet3d.pt", dict(shapes=shapes, seed=91, x_seed=92, ctx_seed=93, t=481, fps=8, y=y))


def g_mvgen():
    """One full MultiViewBaseModel.forward (tiny channels, b=2 CFG halves, m=2 views, F=16) + 2 loop steps."""
    from src.models.MVGenModel import MultiViewBaseModel
    import src.utils.utils as U
    import src.models.MVGenModel as MV
    pers, s_pers = tiny_unet(101)
    pano, s_pano = tiny_unet(102)
    mv = MultiViewBaseModel(pers, pano, pano_pad=True).eval()
    shapes = shapes_of(mv)
    sd = synth_state(shapes, 103)
    missing, unexpected = mv.load_state_dict(sd, strict=False)
    assert not unexpected
    m_, f, b = 2, 16, 2
    cams = tiny_cameras(m_)
    ct = {k: torch.tensor(v)[None] for k, v in cams.items()}        # [b=1, m] as in get_cameras
    lat = synth_tensor((b, m_, 9, f, 16, 16), 104)
    plat = synth_tensor((b, 9, f, 32, 64), 105)
    txt_pers = synth_tensor((b * m_, 5, 32), 106)
    txt_pano = synth_tensor((b, 5, 32), 107)
    feats_pano = synth_tensor((b, f, 4096, 8), 108)
    feats_pers = synth_tensor((b, 1, f, 4096, 8), 109).repeat(1, m_, 1, 1, 1)
    rel = torch.tensor([1.0, 1.0, 63.0, 63.0, 128.0, 256.0])[None, None].repeat(b, f, 1)
    pitch = torch.linspace(-5, 5, f)[None].repeat(b, 1)
    draws = [0.9, 0.1, 0.9, 0.1, 0.9, 0.9, 0.1]
    it = iter(draws)
    U.random.random = lambda: next(it)
    noise_p = synth_tensor((b, 16, 32), 110)
    noise_q = synth_tensor((b * m_, 16, 32), 111)
    nz = iter([noise_p, noise_q])
    MV.torch.randn_like = lambda c: next(nz)
    try:
        ys, yp = mv(latents=lat, pano_latent=plat, timestep=torch.tensor([481]), prompt_embd=txt_pers,
                    pano_prompt_embd=txt_pano, cameras=ct, use_fps_condition=True, use_ip_plus_cross_attention=True,
                    fps_tensor_pano=torch.tensor([8, 8]), fps_tensor_pers=torch.tensor([[8] * m_] * b),
                    reference_images_clip_feat_pano=feats_pano, reference_images_clip_feat_pers=feats_pers,
                    relative_position_tensor=rel, pitchs_tensor=pitch)
    finally:
        import random as _r
        U.random.random = _r.random
        MV.torch.randn_like = torch.randn_like
    save("mvgen.pt", dict(shapes=shapes, seed=103, cams=cams, draws=[d < 0.4 for d in draws], t=481, pers=ys, pano=yp,
                          seeds=dict(lat=104, plat=105, txt_pers=106, txt_pano=107, feats_pano=108, feats_pers=109,
                                     noise_pano=110, noise_pers=111)))


ALL = dict(resnet=g_resnet, transformer=g_transformer, motion=g_motion, geometry=g_geometry, warp=g_warp,
           adapter=g_adapter, ddim=g_ddim, vae=g_vae, unet=g_unet, mvgen=g_mvgen)

if __name__ == "__main__":
    names = sys.argv[1:] or list(ALL)
    for n in names:
        print(f"== {n}")
        ALL[n]()
