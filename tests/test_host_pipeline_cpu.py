"""CPU: the DDIM loop of ``AnimationPipeline.denoise`` (animatediff/pipelines/pipeline_animation_inference_dual.py:734-809) --
latent / mask / masked-latent concatenation per branch, CFG doubling, timestep schedule, the dual-branch step, CFG combine +
DDIM update, resumable ``step_range`` -- executed with every kernel entry point replaced by a torch emulation of its contract
(tests/cpu_ops.py) and compared with the oracle's loop in fp32 on the same bf16-rounded weights, inputs and injected RNG draws
(the comparison tests/test_pipeline_gpu.py::test_denoise_loop_two_steps_vs_oracle makes on the GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cpu_ops import cpu_ops  # noqa: E402
from golden_util import load, synth_state, synth_tensor  # noqa: E402
from test_host_modules import tiny_unet  # noqa: E402
from test_oracle_golden import TINY  # noqa: E402

torch.set_grad_enabled(False)
BF = torch.bfloat16


def test_denoise_loop_host_wiring_vs_oracle(monkeypatch):
    from imagine360_b200.host.config import SCHEDULER_KWARGS
    from imagine360_b200.host.ddim import DDIMScheduler
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    from imagine360_b200.host.pipeline import AnimationPipeline, Conditioning
    from oracle.ddim import DDIM, cfg_combine
    from oracle.mvgen import mv_forward
    monkeypatch.setenv("I360_CUDA_GRAPH", "0")          # eager launch sequence (a CUDA graph needs the device)
    g = load("mvgen.pt")
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet()).to(BF)
    sd = {k: v.to(BF) for k, v in synth_state(g["shapes"], g["seed"]).items()}
    mv.load_state_dict(sd, strict=False)
    sd_o = {k: v.float() for k, v in sd.items()}
    for k, v in mv.state_dict().items():
        if k.endswith(("pos_encoder.pe", "pe.freq_bands")):
            sd_o[k] = v.float()
    f, m = 16, 2
    cams = g["cams"]

    def mk(shape, seed, scale=1.0):
        t = (synth_tensor(shape, seed) * scale).to(BF)
        return t, t.float()

    pano, panoo = mk((1, 4, f, 32, 64), 20)
    pers, perso = mk((1, m, 4, f, 16, 16), 21)
    pmask = torch.ones(1, 1, f, 32, 64)
    pmask[..., 8:24, 16:48] = 0
    vmask = torch.ones(1, m, 1, f, 16, 16)
    pm, pmo = mk((1, 4, f, 32, 64), 22, 0.18)
    vm, vmo = mk((1, m, 4, f, 16, 16), 23, 0.18)
    tp, tpo = mk((2, 5, 32), 24)
    tv, tvo = mk((2 * m, 5, 32), 25)
    fp, fpo = mk((2, f, 4096, 8), 26)
    fv, fvo = mk((2, 1, f, 4096, 8), 27)
    rel = torch.tensor([1.0, 1.0, 63.0, 63.0, 128.0, 256.0])[None].repeat(f, 1)
    pitch = torch.linspace(-5, 5, f)
    draws = [[False, True, False, True, False, False, True], [True, False, False, False, True, False, False]]
    noises = [(mk((2, 16, 32), 30 + i), mk((2 * m, 16, 32), 40 + i)) for i in range(2)]
    pipe = AnimationPipeline(None, None, None, mv.unet, mv.pano_unet, mv, DDIMScheduler(**SCHEDULER_KWARGS))
    cond = Conditioning(tp, tv, fp, fv.expand(-1, m, -1, -1, -1), rel, pitch, 8)
    inject = lambda i: dict(antipodal_draws=draws[i], ip_noise=(noises[i][0][0], noises[i][1][0]))      # noqa: E731
    calls = {}
    with cpu_ops(calls):
        a, b = pipe.denoise(pano, pers, pmask, vmask, pm, vm, cond, cams, 50, 7.5, step_range=(0, 1), inject=inject)
        a, b = pipe.denoise(a, b, pmask, vmask, pm, vm, cond, cams, 50, 7.5, step_range=(1, 2), inject=inject)      # resumed
    assert calls["cfg_ddim_step"] == 4                   # two branches x two steps, one fused CFG + DDIM update each
    sched = DDIM()
    ts = sched.set_timesteps(50)
    pl, vl = panoo, perso
    fps_pano, fps_pers = torch.tensor([8, 8]), torch.tensor([[8] * m] * 2)
    for i in range(2):
        xin_p = torch.cat([pl, pmask, pmo], 1)
        xin_v = torch.cat([vl, vmask, vmo], 2)
        pv, pp = mv_forward(sd_o, torch.cat([xin_v] * 2), torch.cat([xin_p] * 2), ts[i].reshape(1), tvo, tpo, cams, fps_pano, fps_pers,
                            fpo, fvo.expand(-1, m, -1, -1, -1), rel[None].repeat(2, 1, 1), pitch[None].repeat(2, 1), draws[i],
                            noises[i][0][1], noises[i][1][1], cfg=TINY, grid_dtype=BF, pe_dtype=BF)
        pl = sched.step(cfg_combine(pp), int(ts[i]), pl)
        vl = sched.step(cfg_combine(pv), int(ts[i]), vl)
    rel_err = lambda x, y: ((x.float() - y.float()).abs().max() / y.float().abs().max()).item()      # noqa: E731
    e1, e2 = rel_err(a, pl), rel_err(b, vl)
    assert e1 < 6e-2 and e2 < 6e-2, (e1, e2)             # the GPU test's bound


VAE_KW = dict(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
              block_out_channels=(16, 32, 64, 64), layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=8,
              sample_size=64)


def _vae_case(kw, shapes, seed):
    from imagine360_b200.host.vae import AutoencoderKL
    vae = AutoencoderKL(**kw)
    if shapes is None:
        shapes = {k: list(v.shape) for k, v in vae.state_dict().items()}
    assert {k: list(v.shape) for k, v in vae.state_dict().items()} == shapes
    sd = {k: v.to(BF) for k, v in synth_state(shapes, seed).items()}
    vae.load_state_dict(sd)
    return vae.to(BF), {k: v.float() for k, v in sd.items()}


def test_vae_host_wiring_vs_oracle_and_reference():
    """AutoencoderKL.encode / decode (diffusers/models/autoencoder_kl.py; vae.encode at pipeline...dual.py:331-355, decode at
    :301-313): resnets with the conv -> GroupNorm statistics hand-over, asymmetric-pad stride-2 convs, sub-pixel upsample convs,
    single-head mid-block attention as GEMM -> softmax -> GEMM, quant / post-quant convs -- against the oracle and against the
    UNMODIFIED reference's outputs for the same weights (tests/golden/vae.pt)."""
    from oracle import vae as OV
    g = load("vae.pt")
    vae, sd_o = _vae_case(VAE_KW, g["shapes"], g["seed"])
    img = synth_tensor((2, 3, 32, 48), g["img_seed"]).to(BF)
    z = synth_tensor((2, 4, 4, 6), g["z_seed"]).to(BF)
    calls = {}
    with cpu_ops(calls):
        mom = vae.encode(img, 2).latent_dist.parameters
        dec = vae.decode(z).sample
    assert calls["conv3x3_s2"] == 3 and calls["conv_upsample2x"] == 3 and calls["softmax_rows"] >= 2
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()      # noqa: E731
    assert rel(mom, OV.encode_moments(sd_o, img.float(), groups=8)) < 4e-2
    assert rel(dec, OV.decode(sd_o, z.float(), groups=8)) < 4e-2
    assert rel(mom, g["moments"]) < 5e-2 and rel(dec, g["dec"]) < 5e-2          # vs the reference itself


def test_vae_statistics_route_is_the_default_and_optional(monkeypatch):
    """With 32 groups of 4 / 8 / 16 channels (the SD-2.1 widths) the decoder's GroupNorms take their statistics from the producing
    conv's epilogue; with the switch off every GroupNorm runs its own statistics pass.  Same result either way."""
    from imagine360_b200.host import vae as V
    from imagine360_b200.host.config import FULL_VAE_KWARGS
    vae, _ = _vae_case(FULL_VAE_KWARGS, None, 3)
    z = synth_tensor((1, 4, 4, 6), 4).to(BF)
    outs = {}
    for fused in (True, False):
        monkeypatch.setattr(V, "GN_FUSED", fused)
        calls = {}
        with cpu_ops(calls):
            outs[fused] = vae.decode(z).sample
        handed_over = calls.get("groupnorm+stats", 0)
        assert (handed_over > 0) == fused, calls
    assert torch.equal(outs[True], outs[False])      # the emulated statistics are exact either way: the wiring is the same function


def test_init_noise_matches_oracle_exactly():
    """e2p of the panorama noise into the 20 views (pipeline...dual.py:620-640) through the host's batched geometry."""
    from imagine360_b200.host.pipeline import AnimationPipeline
    from oracle import geometry as G
    from oracle import pipeline as OP
    cams = G.default_cameras()
    noise = synth_tensor((1, 4, 1, 4, 32, 64), 5)
    pipe = AnimationPipeline(None, None, None, None, None, None, None)
    with cpu_ops():
        pn, vn = pipe.init_noise(1, 4, 32, 64, 16, 16, cams, "cpu", BF, pano_noise=noise)
    po, vo = OP.init_noise(noise, cams, (16, 16), BF)
    assert torch.equal(pn, po) and torch.equal(vn, vo)


def test_call_end_to_end_host_wiring_vs_oracle(monkeypatch):
    """``AnimationPipeline.__call__`` (pipeline_animation_inference_dual.py:553-815) with stub CLIP / SAM modules at tiny widths:
    init_noise -> VAE encode of the masked panorama and views (chunks of 8, posterior samples) -> text / SAM conditioning -> DDIM
    loop (CFG, the 7 antipodal draws and the two IP-noise draws per step) -> decode_video (pad 4 latent columns -> decode -> crop
    32 px), against the same sequence over the oracle functions.  The RNG order is checked implicitly: the oracle re-seeds torch
    and Python's ``random`` and draws in the reference's order; another order on the native side changes the noise and the
    video by O(1).  (The streamed decode needs CUDA streams and pinned memory; its equality with ``decode_video`` is a GPU test.)"""
    import random
    from golden_util import tiny_cameras
    from imagine360_b200.host.config import SCHEDULER_KWARGS
    from imagine360_b200.host.ddim import DDIMScheduler
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    from imagine360_b200.host.pipeline import AnimationPipeline
    from test_pipeline_call_cpu import StubSam, StubTextEncoder, StubTokenizer
    from test_pipeline_call_gpu import oracle_call
    monkeypatch.setenv("I360_CUDA_GRAPH", "0")
    g = load("mvgen.pt")
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet()).to(BF)
    sd = {k: v.to(BF) for k, v in synth_state(g["shapes"], g["seed"]).items()}
    mv.load_state_dict(sd, strict=False)
    sd_o = {k: v.float() for k, v in sd.items()}
    for k, v in mv.state_dict().items():
        if k.endswith(("pos_encoder.pe", "pe.freq_bands")):
            sd_o[k] = v.float()
    gv = load("vae.pt")
    vae, vsd_o = _vae_case(VAE_KW, gv["shapes"], gv["seed"])
    pipe = AnimationPipeline(vae=vae, text_encoder=StubTextEncoder(32), tokenizer=StubTokenizer(), pers_unet=mv.unet,
                             pano_unet=mv.pano_unet, mv_base_model=mv, scheduler=DDIMScheduler(**SCHEDULER_KWARGS),
                             image_encoder=StubSam(img_size=1024), image_encoder_name="SAM")
    pipe.enable_vae_slicing()
    monkeypatch.setattr(pipe, "decode_video_streamed", lambda lat: (pipe.decode_video(lat).cpu(), None), raising=False)
    f, m, H, W, ps = 16, 2, 256, 512, 128
    pano_mask = torch.ones(1, f, 1, H, W)
    pano_mask[..., H // 4: 3 * H // 4, W // 2 - H // 4: W // 2 + H // 4] = 0
    pers_masks = torch.ones(1, f, m, 1, ps, ps)
    pers_masks[:, :, 0, :, 16:112, 16:112] = 0
    vb = {"fps": 8, "video_length": f, "pano_H": H, "pano_W": W, "pers_size": ps, "cameras": tiny_cameras(m),
          "pano_pixel_values": synth_tensor((1, f, 3, H, W), 50, 0.5).clamp(-1, 1).to(BF), "pano_mask": pano_mask.to(BF),
          "pers_pixel_values": synth_tensor((1, f, m, 3, ps, ps), 51, 0.5).clamp(-1, 1).to(BF), "pers_masks": pers_masks.to(BF),
          "anchor_pixels_values": synth_tensor((1, f, 3, 63, 63), 52, 0.5).clamp(-1, 1).to(BF),
          "anchor_pixels_values_pers": synth_tensor((1, f, 3, 48, 64), 53, 0.5).clamp(-1, 1).to(BF),
          "relative_position": torch.tensor([[1, 1, 63, 63, H, W]] * f).to(BF), "pitchs": torch.linspace(-4, 4, f).to(BF)}
    steps, seed_t, seed_p = 2, 1234, 77
    torch.manual_seed(seed_t)
    random.seed(seed_p)
    with cpu_ops():
        out = pipe("a tiny panorama", latents_dtype=BF, video_batch=vb, num_inference_steps=steps, use_outpaint=True,
                   use_ip_plus_cross_attention=True, ip_plus_condition="video", use_fps_condition=True, negative_prompt="blurry").videos
    assert out.shape == (1, 3, f, H, W) and out.dtype == torch.float32
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    ref = oracle_call(sd_o, vsd_o, pipe, vb, "a tiny panorama", "blurry", steps, torch.float32, seed_t, seed_p, dev="cpu")
    d = (out - ref).abs()
    # bf16 storage between the ops of two steps and a decode against fp32 throughout: measured mean 0.008, max 0.09 of a [0, 1] video;
    # a wrong RNG order, chunking, mask or crop moves the video by O(0.3)
    assert d.mean().item() < 0.02 and d.max().item() < 0.2, (d.mean().item(), d.max().item())


def test_step_graph_construction_logic(monkeypatch):
    """``StepGraph.__init__`` -- static buffers, conditioning load, the two eager warm-up passes (both WarpAttn bias variants of all
    7 call sites), bias slots frozen, adapter + step captures -- executed on the CPU with torch's CUDA-graph API stubbed out and
    the kernels emulated: checks the PYTHON of the default GPU path (the graphs themselves are a GPU test:
    tests/test_pipeline_gpu.py::test_cuda_graph_step_equals_eager_step).  The graph object holds the geometry tables it reads by
    address, so pruning the geometry cache cannot free them."""
    import contextlib
    import types
    from imagine360_b200.host import geometry as G
    from imagine360_b200.host import pipeline as PL
    from imagine360_b200.host.config import SCHEDULER_KWARGS
    from imagine360_b200.host.ddim import DDIMScheduler
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    monkeypatch.setattr(torch.cuda, "CUDAGraph", lambda: types.SimpleNamespace(replay=lambda: None))
    monkeypatch.setattr(torch.cuda, "graph", lambda g: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    G._grid_cache.clear()
    G._set_last_use.clear()
    g = load("mvgen.pt")
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet()).to(BF)
    mv.load_state_dict({k: v.to(BF) for k, v in synth_state(g["shapes"], g["seed"]).items()}, strict=False)
    f, m = 16, 2
    mk = lambda shape, seed, scale=1.0: (synth_tensor(shape, seed) * scale).to(BF)      # noqa: E731
    pano, pers = mk((1, 4, f, 32, 64), 20), mk((1, m, 4, f, 16, 16), 21)
    cond = PL.Conditioning(mk((2, 5, 32), 24), mk((2 * m, 5, 32), 25), mk((2, f, 4096, 8), 26),
                           mk((2, 1, f, 4096, 8), 27).expand(-1, m, -1, -1, -1),
                           torch.tensor([1.0, 1.0, 63.0, 63.0, 128.0, 256.0])[None].repeat(f, 1), torch.linspace(-5, 5, f), 8)
    pipe = PL.AnimationPipeline(None, None, None, mv.unet, mv.pano_unet, mv, DDIMScheduler(**SCHEDULER_KWARGS))
    calls = {}
    with cpu_ops(calls):
        sg = PL.StepGraph(pipe, pano, pers, cond, g["cams"])
    assert sg.pred_pano.shape == (2, 4, f, 32, 64) and sg.pred_pers.shape == (2, m, 4, f, 16, 16)
    assert len(sg.slots) == 7 and calls["grid_sample"] > 0
    held = sg._geometry
    assert len(held) > 0 and len(held) == len(G.cached_entries(sg.cams))
    G._grid_cache.clear()                      # an aggressive prune: the graph's tables stay referenced
    assert all(v is not None for v in held)
    G._set_last_use.clear()
