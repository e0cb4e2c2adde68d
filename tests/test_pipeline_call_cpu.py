"""CPU checks of the drop-in boundary of ``AnimationPipeline``: the constructor wraps a RAW SAM model the
way the reference does (pipeline_animation_inference_dual.py:171-174), ``__call__`` accepts exactly the kwargs and
``video_batch`` keys ``inference_dual_p2e.py:548-564,:584-595`` passes, the SAM feature path follows :675-718, and the
conditioning reaches ``denoise`` with the reference's shapes.  The kernel-backed stages (init_noise, VAE, the loop, the
decode) are replaced by recorders here -- they are covered on the GPU by tests/test_pipeline_call_gpu.py."""
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from imagine360_b200.host import pipeline as PL  # noqa: E402

BF = torch.bfloat16


class StubSamEncoder(nn.Module):
    """Stands in for SAM's ImageEncoderViT: [B, 3, S, S] -> [B, C, S/16, S/16]."""

    def __init__(self, c=8, img_size=64):
        super().__init__()
        self.img_size = img_size
        torch.manual_seed(3)
        self.proj = nn.Conv2d(3, c, 16, stride=16)

    def forward(self, x):
        assert x.shape[-2:] == (self.img_size, self.img_size) and x.dtype == torch.float32
        return self.proj(x)


class StubSam(nn.Module):
    """What ``sam_model_registry['vit_b']()`` returns, reduced to the members SamPredictor touches."""

    def __init__(self, img_size=64):
        super().__init__()
        self.image_encoder = StubSamEncoder(img_size=img_size)
        self.register_buffer("pixel_mean", torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1), False)

    def preprocess(self, x):
        x = (x - self.pixel_mean) / self.pixel_std
        s = self.image_encoder.img_size
        return torch.nn.functional.pad(x, (0, s - x.shape[-1], 0, s - x.shape[-2]))


class StubTokenizer:
    model_max_length = 7

    def __call__(self, texts, padding, max_length, truncation, return_tensors):
        ids = torch.tensor([[(sum(map(ord, t)) + i) % 50 for i in range(max_length)] for t in texts])
        return SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))


class StubTextEncoder(nn.Module):
    def __init__(self, d=12):
        super().__init__()
        torch.manual_seed(4)
        self.emb = nn.Embedding(50, d)
        self.config = SimpleNamespace(use_attention_mask=False)

    def forward(self, ids, attention_mask=None):
        assert attention_mask is None
        return (self.emb(ids),)


def reference_sam_features(sam, anchor):
    """pipeline...dual.py:679-694 written out literally (numpy uint8 conversion, apply_image per frame through
    torchvision exactly as segment_anything's ResizeLongestSide does, batches of 8, 'f c h w -> f (h w) c')."""
    from torchvision.transforms.functional import resize, to_pil_image
    image_array = np.uint8(((anchor.to(torch.float32) + 1.0) / 2.0 * 255).cpu().numpy().transpose(0, 2, 3, 1))
    size = sam.image_encoder.img_size
    tens = []
    for image in image_array:
        oldh, oldw = image.shape[:2]
        scale = size * 1.0 / max(oldh, oldw)
        target = (int(oldh * scale + 0.5), int(oldw * scale + 0.5))
        tens.append(torch.as_tensor(np.array(resize(to_pil_image(image), target))).permute(2, 0, 1).contiguous())
    tens = torch.stack(tens)
    out = []
    for i in range(tens.shape[0] // 8):
        feats = sam.image_encoder(sam.preprocess(tens[i * 8:(i + 1) * 8]))
        out.append(feats.flatten(2).transpose(1, 2))
    return torch.cat(out)


def make_pipeline():
    vae = SimpleNamespace(config=SimpleNamespace(block_out_channels=(1, 2, 3, 4)), dtype=BF)
    # keyword construction exactly as inference_dual_p2e.py:463-473
    return PL.AnimationPipeline(pers_unet=None, pano_unet=None, mv_base_model=nn.Identity(), vae=vae, tokenizer=StubTokenizer(),
                                text_encoder=StubTextEncoder(), scheduler=None, image_encoder=StubSam(), image_encoder_name="SAM")


def test_constructor_wraps_raw_sam_model_like_the_reference():
    pipe = make_pipeline()
    assert hasattr(pipe, "SAMpredictor") and pipe.SAMProcessor is pipe.SAMpredictor.transform
    assert pipe.vae_scale_factor == 8
    img = (np.random.default_rng(0).random((37, 50, 3)) * 255).astype(np.uint8)
    from torchvision.transforms.functional import resize, to_pil_image
    want = np.array(resize(to_pil_image(img), (47, 64)))
    assert np.array_equal(pipe.SAMProcessor.apply_image(img), want)


def test_sam_features_follow_the_reference_sequence():
    pipe = make_pipeline()
    torch.manual_seed(0)
    anchor = (torch.rand(16, 3, 31, 31) * 2 - 1).to(BF)          # anchor crops are odd-sized (127 x 127 in production)
    got = pipe._sam_features(anchor)
    want = reference_sam_features(pipe.image_encoder, anchor)
    assert got.shape == (16, 16, 8) and torch.equal(got, want)
    with pytest.raises(ValueError):
        pipe._sam_features(anchor[:12])                          # F % 8 != 0 (pipeline...dual.py:685)


def test_call_accepts_the_scripts_kwargs_and_reaches_denoise(monkeypatch):
    pipe = make_pipeline().to("cpu")
    pipe.enable_vae_slicing = lambda: None
    f, m, H, W, ps = 16, 3, 64, 128, 32
    seen = {}

    def fake_init_noise(bs, video_length, eh, ew, ph, pw, cameras, device, dtype):
        seen["init_noise"] = (bs, video_length, eh, ew, ph, pw)
        return torch.zeros(1, 4, f, eh, ew, dtype=dtype), torch.zeros(1, m, 4, f, ph, pw, dtype=dtype)

    def fake_denoise(pano_latent, pers_latent, pano_mask, pers_masks, pano_masked, pers_masked, cond, cameras, steps, guidance):
        seen["denoise"] = dict(cond=cond, steps=steps, guidance=guidance, pano_mask=pano_mask.shape, pers_masks=pers_masks.shape)
        return pano_latent, pers_latent

    monkeypatch.setattr(pipe, "init_noise", fake_init_noise)
    monkeypatch.setattr(pipe, "prepare_masked_latents_pano", lambda f_, px, mask: (torch.zeros(1, 4, f, H // 8, W // 8), torch.zeros(1, 1, f, H // 8, W // 8)))
    monkeypatch.setattr(pipe, "prepare_masked_latents_pers", lambda f_, px, mask: (torch.zeros(1, m, 4, f, ps // 8, ps // 8), torch.zeros(1, m, 1, f, ps // 8, ps // 8)))
    monkeypatch.setattr(pipe, "denoise", fake_denoise)
    monkeypatch.setattr(pipe, "decode_video_streamed", lambda lat: (torch.zeros(1, 3, f, H, W), torch.zeros(f, H, W, 3, dtype=torch.uint8)))
    g = torch.Generator().manual_seed(1)
    vb = {                                                        # inference_dual_p2e.py:548-564
        "videoid": "x.mp4", "fps": 8,
        "anchor_pixels_values_pers": (torch.rand(1, f, 3, 24, 40, generator=g) * 2 - 1).to(BF),
        "pano_pixel_values": torch.zeros(1, f, 3, H, W, dtype=BF), "pano_mask": torch.ones(1, f, 1, H, W, dtype=BF),
        "video_length": f, "anchor_pixels_values": (torch.rand(1, f, 3, 31, 31, generator=g) * 2 - 1).to(BF),
        "relative_position": torch.tensor([[1, 1, 31, 31, H, W]] * f).to(BF), "pitchs": torch.zeros(f, dtype=BF),
        "pers_pixel_values": torch.zeros(1, f, m, 3, ps, ps, dtype=BF), "pers_masks": torch.ones(1, f, m, 1, ps, ps, dtype=BF),
        "cameras": {"FoV": torch.full((1, m), 90), "theta": torch.zeros(1, m), "phi": torch.zeros(1, m)},
        "pano_H": H, "pano_W": W, "pers_size": ps,
    }
    out = pipe("a prompt", latents_dtype=BF, video_batch=vb, num_inference_steps=50, use_outpaint=True,
               generator=torch.Generator().manual_seed(0), use_ip_plus_cross_attention=True, ip_plus_condition="video",
               use_fps_condition=True, negative_prompt="bad").videos      # :584-595
    assert out.shape == (1, 3, f, H, W) and out._i360_frames_u8.shape == (f, H, W, 3)
    assert seen["init_noise"] == (1, f, H // 8, W // 8, ps // 8, ps // 8)
    d = seen["denoise"]
    c = d["cond"]
    assert d["steps"] == 50 and d["guidance"] == 7.5
    assert c.text_pano.shape == (2, 7, 12) and c.text_pers.shape == (2 * m, 7, 12) and c.text_pano.dtype == BF
    assert torch.equal(c.text_pers[:m], c.text_pano[:1].expand(m, -1, -1))              # uncond rows first, then cond
    assert c.feats_pano.shape == (2, f, 16, 8) and torch.equal(c.feats_pano[0], c.feats_pano[1])     # cond == uncond copy (:695)
    assert c.feats_pers.shape == (2, m, f, 16, 8) and torch.equal(c.feats_pers[:, 0], c.feats_pers[:, m - 1])
    want = reference_sam_features(pipe.image_encoder, vb["anchor_pixels_values"][0]).to(BF)
    assert torch.equal(c.feats_pano[0], want)
    assert c.rel_pos.shape == (f, 6) and c.pitch.shape == (f,) and c.fps == 8


def test_unsupported_configuration_is_reported_on_stderr(capsys):
    pipe = make_pipeline()
    with pytest.raises(NotImplementedError):
        pipe("p", latents_dtype=BF, video_batch={}, use_outpaint=False, use_ip_plus_cross_attention=True, ip_plus_condition="video",
             use_fps_condition=True)
    assert "imagine360_b200" in capsys.readouterr().err      # the script's bare `except: continue` must not hide it


def test_adapter_cache_does_not_survive_a_new_clip():
    """the adapter cache was keyed on data_ptr/_version only; a freed-and-reallocated feature tensor with
    new contents reproduced the key.  The entry now pins the keyed tensors, so a new clip can never alias an old key."""
    from imagine360_b200.host import mvgen as MV
    calls = []

    class FakeUnet(nn.Module):
        use_relative_postions = False

    mv = MV.MultiViewBaseModel.__new__(MV.MultiViewBaseModel)
    nn.Module.__init__(mv)
    mv.unet, mv.pano_unet, mv._adapter_cache = FakeUnet(), FakeUnet(), {}
    orig = MV.ip_tokens_clean
    MV.ip_tokens_clean = lambda unet, feats: (calls.append(float(feats.sum())), feats.reshape(feats.shape[0], -1)[:, :4].clone())[1]
    try:
        rel, pitch = torch.zeros(2, 16, 6), torch.zeros(2, 16)
        ptrs = set()
        for clip in range(4):
            fp = torch.full((2, 16, 8, 4), float(clip))
            fv = torch.full((2, 1, 16, 8, 4), float(clip)).expand(-1, 3, -1, -1, -1)
            ptrs.add(fp.data_ptr())
            a = mv._adapter(fp, fv, rel, pitch)
            b = mv._adapter(fp, fv.expand(-1, 3, -1, -1, -1), rel, pitch)      # same clip, a fresh view object: cache hit
            assert a[0] is b[0]
            assert float(a[0][0, 0]) == float(clip)
            del fp, fv                                                          # freed: the allocator may reuse the block
        assert len(calls) == 8                                                  # 2 branches x 4 clips, never a stale hit
    finally:
        MV.ip_tokens_clean = orig


def test_data_edit_invalidates_packed_weights():
    """``weight.data += ...`` (the reference's LoRA merge, inference_dual_p2e.py:193) does not bump
    ``_version``; refresh_packed_weights() (run once per clip by the pipeline) catches it by content."""
    from imagine360_b200.host.unet3d import invalidate_packed_weights, lin_w, refresh_packed_weights
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(8, 8), nn.Linear(8, 4))
    assert refresh_packed_weights(net) is False
    w0 = lin_w(net[0])[0]
    v = net[0].weight._version
    net[0].weight.data += 1.0
    assert net[0].weight._version == v                     # the trap
    assert lin_w(net[0])[0] is w0                          # stale without the content check
    assert refresh_packed_weights(net) is True
    w1 = lin_w(net[0])[0]
    assert w1 is not w0 and torch.equal(w1, net[0].weight.to(BF))
    assert refresh_packed_weights(net) is False
    invalidate_packed_weights(net)
    assert lin_w(net[0])[0] is not w1


def test_constructor_options_the_native_path_does_not_implement_raise():
    from test_host_modules import TINY_KW
    from imagine360_b200.host.unet3d import UNet3DConditionModel, VanillaTemporalModule
    with pytest.raises(NotImplementedError):
        UNet3DConditionModel(**{**TINY_KW, "use_inflated_groupnorm": False})
    with pytest.raises(NotImplementedError):
        VanillaTemporalModule(in_channels=32, temporal_position_encoding=False)


def test_scalar_camera_arguments_broadcast():
    from imagine360_b200.host.geometry import camera_lists
    fov, th, ph = camera_lists({"FoV": 90, "theta": [0.0, 10.0, 20.0], "phi": torch.tensor([1.0, 2.0, 3.0])})
    assert fov == (90.0, 90.0, 90.0) and th == (0.0, 10.0, 20.0) and ph == (1.0, 2.0, 3.0)
