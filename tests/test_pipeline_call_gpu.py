"""End-to-end ``AnimationPipeline.__call__`` on the GPU (tiny widths, stub CLIP / SAM -- third-party encoders, out of scope)
against ``oracle/pipeline.py``: init_noise -> VAE encode of the masked panorama and the masked views (chunks of 8,
posterior samples) -> text / SAM conditioning -> DDIM loop (CFG, the 7 antipodal draws and the two IP-noise draws per step)
-> ``decode_video`` (pad 4 latent columns -> decode -> crop 32 px) -> fp32 CPU video in [0, 1].
pipeline_animation_inference_dual.py:553-815.

RNG order (SURVEY.md trap 3) is checked implicitly: the oracle re-seeds torch's CUDA generator and Python's ``random``
and draws the same shapes / dtypes in the reference's order; a different order on the native side changes the noise and the
comparison fails by O(1)."""
import os
import random
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_util import load, synth_state, synth_tensor, tiny_cameras  # noqa: E402
from test_blocks_gpu import BF, TINY, load_native, q  # noqa: E402
from test_host_modules import tiny_unet  # noqa: E402
from test_parity_calibrated_gpu import calibrated  # noqa: E402
from test_pipeline_call_cpu import StubSam, StubTextEncoder, StubTokenizer  # noqa: E402
from test_pipeline_gpu import VAE_KW  # noqa: E402

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def sam_1024():
    """SAM ViT-B geometry (1024-px input, 64 x 64 = 4096 tokens) with a one-conv encoder of 8 channels."""
    return StubSam(img_size=1024)


def oracle_call(sd_mv, sd_vae, pipe, vb, prompt, negative, steps, dtype, seed_t, seed_p, dev="cuda"):
    """The reference's __call__ sequence over the oracle functions, in ``dtype`` (fp32 = truth with the production grid /
    PE quantisation; bf16 = what the reference's production path computes through torch's kernels)."""
    from oracle import pipeline as OP
    from oracle import vae as OV
    f, m = vb["video_length"], vb["pers_pixel_values"].shape[2]
    cams = vb["cameras"]
    torch.manual_seed(seed_t)
    random.seed(seed_p)
    pano_px = (vb["pano_pixel_values"].clone() * (vb["pano_mask"] < 0.5)).to(dev)
    pers_px = (vb["pers_pixel_values"].clone() * (vb["pers_masks"] < 0.5)).to(dev)
    eh, ew, ph = vb["pano_H"] // 8, vb["pano_W"] // 8, vb["pers_size"] // 8
    noise = torch.randn(1, f, 1, 4, eh, ew, device=dev)
    pano_lat, pers_lat = OP.init_noise(noise, cams, (ph, ph), BF)

    def encode(x):          # chunks of 8, moments -> posterior sample with a fresh fp32 randn (vae.py:354-361)
        out = []
        for i in range(0, x.shape[0], 8):
            mom = OV.encode_moments(sd_vae, x[i:i + 8].to(dtype), groups=8)
            out.append(OV.sample_posterior(mom, torch.randn(mom.shape[0], 4, *mom.shape[2:], device=dev)))
        return torch.cat(out)

    lat = encode(pano_px.reshape(-1, *pano_px.shape[2:]))
    pano_masked = lat.reshape(1, f, *lat.shape[1:]).permute(0, 2, 1, 3, 4) * 0.18215
    mask = vb["pano_mask"].to(dev).transpose(2, 1)
    pano_mask = F.interpolate(mask, size=(mask.shape[2], eh, ew))
    lat = encode(pers_px.reshape(-1, *pers_px.shape[3:]))
    pers_masked = lat.reshape(1, f, m, *lat.shape[1:]).permute(0, 2, 3, 1, 4, 5) * 0.18215
    masks = vb["pers_masks"].to(dev).permute(0, 3, 1, 2, 4, 5).squeeze(0)
    masks = F.interpolate(masks, size=(m, ph, ph)).unsqueeze(3).permute(0, 2, 3, 1, 4, 5)
    text_pano = pipe._encode_prompt([prompt], dev, [negative]).to(BF)
    text_pers = pipe._encode_prompt([prompt] * m, dev, [negative] * m).to(BF)
    feats = pipe._sam_features(vb["anchor_pixels_values"].to(dev)[0]).to(BF)[None]
    feats_p = pipe._sam_features(vb["anchor_pixels_values_pers"].to(dev)[0]).to(BF)[None]
    cond = dict(text_pano=text_pano.to(dtype), text_pers=text_pers.to(dtype), feats_pano=torch.cat([feats, feats]).to(dtype),
                feats_pers=torch.cat([feats_p, feats_p]).unsqueeze(1).expand(-1, m, -1, -1, -1).to(dtype), fps=vb["fps"],
                rel_pos=vb["relative_position"].to(dev).reshape(f, 6).to(dtype), pitch=vb["pitchs"].to(dev).reshape(f).to(dtype))
    kw = dict(grid_dtype=BF, pe_dtype=BF) if dtype == torch.float32 else {}
    pano_lat, _ = OP.denoise_loop(sd_mv, pano_lat.to(dtype), pers_lat.to(dtype), pano_mask.to(dtype), masks.to(dtype),
                                  pano_masked.to(dtype), pers_masked.to(dtype), cond, cams, steps, 7.5, cfg=TINY,
                                  noise_fn=lambda shape: torch.randn(shape, dtype=BF, device=dev).to(dtype), **kw)
    return OP.decode_video(sd_vae, pano_lat, groups=8).cpu()


def test_call_end_to_end_vs_oracle():
    from imagine360_b200.host.config import SCHEDULER_KWARGS
    from imagine360_b200.host.ddim import DDIMScheduler
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    from imagine360_b200.host.pipeline import AnimationPipeline
    from imagine360_b200.host.vae import AutoencoderKL
    g = load("mvgen.pt")
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet())
    sd_n, sd_o = q(synth_state(g["shapes"], g["seed"]))
    load_native(mv, sd_n)
    for k in g["shapes"]:
        if k.endswith(("pos_encoder.pe", "pe.freq_bands")):
            sd_o[k] = mv.state_dict()[k].float()
            sd_n[k] = mv.state_dict()[k]
    gv = load("vae.pt")
    vae = AutoencoderKL(**VAE_KW)
    vsd_n, vsd_o = q(synth_state(gv["shapes"], gv["seed"]))
    load_native(vae, vsd_n)
    pipe = AnimationPipeline(vae=vae, text_encoder=StubTextEncoder(32).cuda(), tokenizer=StubTokenizer(), pers_unet=mv.unet,
                             pano_unet=mv.pano_unet, mv_base_model=mv, scheduler=DDIMScheduler(**SCHEDULER_KWARGS),
                             image_encoder=sam_1024().cuda(), image_encoder_name="SAM").to("cuda")
    pipe.enable_vae_slicing()
    assert pipe.vae_scale_factor == 8
    f, m, H, W, ps = 16, 2, 256, 512, 128
    cams = tiny_cameras(m)
    pano_mask = torch.ones(1, f, 1, H, W)
    pano_mask[..., H // 4: 3 * H // 4, W // 2 - H // 4: W // 2 + H // 4] = 0
    pers_masks = torch.ones(1, f, m, 1, ps, ps)
    pers_masks[:, :, 0, :, 16:112, 16:112] = 0
    vb = {"fps": 8, "video_length": f, "pano_H": H, "pano_W": W, "pers_size": ps, "cameras": cams,
          "pano_pixel_values": synth_tensor((1, f, 3, H, W), 50, 0.5).clamp(-1, 1).to(BF), "pano_mask": pano_mask.to(BF),
          "pers_pixel_values": synth_tensor((1, f, m, 3, ps, ps), 51, 0.5).clamp(-1, 1).to(BF), "pers_masks": pers_masks.to(BF),
          "anchor_pixels_values": synth_tensor((1, f, 3, 63, 63), 52, 0.5).clamp(-1, 1).to(BF),
          "anchor_pixels_values_pers": synth_tensor((1, f, 3, 48, 64), 53, 0.5).clamp(-1, 1).to(BF),
          "relative_position": torch.tensor([[1, 1, 63, 63, H, W]] * f).to(BF), "pitchs": torch.linspace(-4, 4, f).to(BF)}
    steps, seed_t, seed_p = 3, 1234, 77
    torch.manual_seed(seed_t)
    random.seed(seed_p)
    out = pipe("a tiny panorama", latents_dtype=BF, video_batch=vb, num_inference_steps=steps, use_outpaint=True,
               generator=torch.Generator(device="cuda").manual_seed(0), use_ip_plus_cross_attention=True,
               ip_plus_condition="video", use_fps_condition=True, negative_prompt="blurry").videos
    assert out.shape == (1, 3, f, H, W) and out.dtype == torch.float32 and out.device.type == "cpu"
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    ref32 = oracle_call(sd_o, vsd_o, pipe, vb, "a tiny panorama", "blurry", steps, torch.float32, seed_t, seed_p)
    ref16 = oracle_call(sd_n, vsd_n, pipe, vb, "a tiny panorama", "blurry", steps, BF, seed_t, seed_p)
    calibrated(out, ref32, ref16, "AnimationPipeline.__call__ video (3 DDIM steps)")
    # a second clip through the same pipeline object: new conditioning, caches must not leak
    vb2 = dict(vb)
    vb2["anchor_pixels_values"] = synth_tensor((1, f, 3, 63, 63), 60, 0.5).clamp(-1, 1).to(BF)
    vb2["anchor_pixels_values_pers"] = synth_tensor((1, f, 3, 48, 64), 61, 0.5).clamp(-1, 1).to(BF)
    torch.manual_seed(seed_t)
    random.seed(seed_p)
    out2 = pipe("a tiny panorama", latents_dtype=BF, video_batch=vb2, num_inference_steps=steps, use_outpaint=True,
                use_ip_plus_cross_attention=True, ip_plus_condition="video", use_fps_condition=True, negative_prompt="blurry").videos
    ref32b = oracle_call(sd_o, vsd_o, pipe, vb2, "a tiny panorama", "blurry", steps, torch.float32, seed_t, seed_p)
    ref16b = oracle_call(sd_n, vsd_n, pipe, vb2, "a tiny panorama", "blurry", steps, BF, seed_t, seed_p)
    calibrated(out2, ref32b, ref16b, "second clip through the same pipeline")
    assert (out2 - out).abs().max() > 1e-3        # the image prompt does influence the result
