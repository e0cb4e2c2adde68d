"""GPU parity of the VAE, init_noise and the DDIM loop (through the host mirror / C ABI) against the oracle."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_util import load, synth_state, synth_tensor, tiny_cameras  # noqa: E402
from test_blocks_gpu import BF, TINY, load_native, q, qt, rel_err  # noqa: E402
from test_host_modules import tiny_unet  # noqa: E402

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

VAE_KW = dict(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
              block_out_channels=(16, 32, 64, 64), layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=8,
              sample_size=64)


def test_vae_state_dict_and_parity():
    from imagine360_b200.host.vae import AutoencoderKL
    from oracle import vae as OV
    g = load("vae.pt")
    vae = AutoencoderKL(**VAE_KW)
    assert {k: list(v.shape) for k, v in vae.state_dict().items()} == g["shapes"]
    sd_n, sd_o = q(synth_state(g["shapes"], g["seed"]))
    load_native(vae, sd_n)
    img, imgo = qt(synth_tensor((2, 3, 32, 48), g["img_seed"]))
    z, zo = qt(synth_tensor((2, 4, 4, 6), g["z_seed"]))
    mom = vae.encode(img, 2).latent_dist.parameters
    dec = vae.decode(z).sample
    assert rel_err(mom, OV.encode_moments(sd_o, imgo, groups=8)) < 4e-2
    assert rel_err(dec, OV.decode(sd_o, zo, groups=8)) < 4e-2
    assert rel_err(mom, g["moments"].cuda()) < 5e-2 and rel_err(dec, g["dec"].cuda()) < 5e-2   # vs the reference itself


def test_vae_full_width_decode_block():
    """SD-2.1 widths (128..512 channels, 32 groups) on a small latent: exercises the head_dim-512 attention path."""
    from imagine360_b200.host.config import FULL_VAE_KWARGS
    from imagine360_b200.host.vae import AutoencoderKL
    from oracle import vae as OV
    vae = AutoencoderKL(**FULL_VAE_KWARGS)
    shapes = {k: list(v.shape) for k, v in vae.state_dict().items()}
    sd_n, sd_o = q(synth_state(shapes, 3))
    load_native(vae, sd_n)
    z, zo = qt(synth_tensor((1, 4, 8, 20), 4))
    assert rel_err(vae.decode(z).sample, OV.decode(sd_o, zo)) < 5e-2


def test_init_noise_matches_oracle_exactly():
    from imagine360_b200.host.pipeline import AnimationPipeline
    from oracle import pipeline as OP
    from oracle import geometry as G
    cams = G.default_cameras()
    noise = synth_tensor((1, 4, 1, 4, 32, 64), 5)
    pipe = AnimationPipeline(None, None, None, None, None, None, None)
    pn, vn = pipe.init_noise(1, 4, 32, 64, 16, 16, cams, "cuda", BF, pano_noise=noise.cuda())
    po, vo = OP.init_noise(noise, cams, (16, 16), BF)
    assert torch.equal(pn.cpu(), po) and torch.equal(vn.cpu(), vo)


def test_denoise_loop_two_steps_vs_oracle():
    from imagine360_b200.host.config import SCHEDULER_KWARGS
    from imagine360_b200.host.ddim import DDIMScheduler
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    from imagine360_b200.host.pipeline import AnimationPipeline, Conditioning
    from oracle import pipeline as OP
    g = load("mvgen.pt")
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet())
    sd_n, sd_o = q(synth_state(g["shapes"], g["seed"]))
    load_native(mv, sd_n)
    for k in g["shapes"]:
        if k.endswith(("pos_encoder.pe", "pe.freq_bands")):
            sd_o[k] = mv.state_dict()[k].float()
    f, m = 16, 2
    cams = g["cams"]

    def mk(shape, seed, scale=1.0):
        return qt(synth_tensor(shape, seed) * scale)

    pano, panoo = mk((1, 4, f, 32, 64), 20)
    pers, perso = mk((1, m, 4, f, 16, 16), 21)
    pmask = torch.ones(1, 1, f, 32, 64).cuda()
    pmask[..., 8:24, 16:48] = 0
    vmask = torch.ones(1, m, 1, f, 16, 16).cuda()
    pm, pmo = mk((1, 4, f, 32, 64), 22, 0.18)
    vm, vmo = mk((1, m, 4, f, 16, 16), 23, 0.18)
    tp, tpo = mk((2, 5, 32), 24)
    tv, tvo = mk((2 * m, 5, 32), 25)
    fp, fpo = mk((2, f, 4096, 8), 26)
    fv, fvo = mk((2, 1, f, 4096, 8), 27)
    rel = torch.tensor([1.0, 1.0, 63.0, 63.0, 128.0, 256.0])[None].repeat(f, 1).cuda()
    pitch = torch.linspace(-5, 5, f).cuda()
    draws = [[False, True, False, True, False, False, True], [True, False, False, False, True, False, False]]
    noises = [(mk((2, 16, 32), 30 + i), mk((2 * m, 16, 32), 40 + i)) for i in range(2)]
    pipe = AnimationPipeline(None, None, None, mv.unet, mv.pano_unet, mv, DDIMScheduler(**SCHEDULER_KWARGS))
    cond = Conditioning(tp, tv, fp, fv.expand(-1, m, -1, -1, -1), rel, pitch, 8)
    a, b = pipe.denoise(pano, pers, pmask, vmask, pm, vm, cond, cams, 50, 7.5, step_range=(0, 2),
                        inject=lambda i: dict(antipodal_draws=draws[i], ip_noise=(noises[i][0][0], noises[i][1][0])))
    # oracle: same two steps in fp32 with the production grid / PE quantisation
    import oracle.pipeline as OPm
    from oracle.ddim import DDIM, cfg_combine
    from oracle.mvgen import mv_forward
    sched = DDIM()
    ts = sched.set_timesteps(50)
    pl, vl = panoo, perso
    fps_pano, fps_pers = torch.tensor([8, 8]).cuda(), torch.tensor([[8] * m] * 2).cuda()
    for i in range(2):
        xin_p = torch.cat([pl, pmask, pmo], 1)
        xin_v = torch.cat([vl, vmask, vmo], 2)
        pv, pp = mv_forward(sd_o, torch.cat([xin_v] * 2), torch.cat([xin_p] * 2), ts[i].reshape(1).cuda(), tvo, tpo, cams, fps_pano, fps_pers,
                            fpo, fvo.expand(-1, m, -1, -1, -1), rel[None].repeat(2, 1, 1), pitch[None].repeat(2, 1), draws[i],
                            noises[i][0][1], noises[i][1][1], cfg=TINY, grid_dtype=BF, pe_dtype=BF)
        pl = sched.step(cfg_combine(pp), int(ts[i]), pl)
        vl = sched.step(cfg_combine(pv), int(ts[i]), vl)
    e1, e2 = rel_err(a, pl), rel_err(b, vl)
    print("denoise 2 steps rel err", e1, e2)
    assert e1 < 6e-2 and e2 < 6e-2


def test_cuda_graph_step_equals_eager_step(monkeypatch):
    """The graphed loop (StepGraph: adapter graph + step graph, bias slots, in-place DDIM) reproduces the eager loop
    bit for bit: same kernels, same torch-RNG stream (IP noise, pano first), same Python ``random`` draws."""
    import random
    from imagine360_b200.host.config import SCHEDULER_KWARGS
    from imagine360_b200.host.ddim import DDIMScheduler
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    from imagine360_b200.host.pipeline import AnimationPipeline, Conditioning
    g = load("mvgen.pt")
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet())
    sd_n, _ = q(synth_state(g["shapes"], g["seed"]))
    load_native(mv, sd_n)
    f, m = 16, 2

    def mk(shape, seed, scale=1.0):
        return qt(synth_tensor(shape, seed) * scale)[0]

    pano, pers = mk((1, 4, f, 32, 64), 20), mk((1, m, 4, f, 16, 16), 21)
    pmask = torch.ones(1, 1, f, 32, 64).cuda()
    pmask[..., 8:24, 16:48] = 0
    vmask = torch.ones(1, m, 1, f, 16, 16).cuda()
    pm, vm = mk((1, 4, f, 32, 64), 22, 0.18), mk((1, m, 4, f, 16, 16), 23, 0.18)
    rel = torch.tensor([1.0, 1.0, 63.0, 63.0, 128.0, 256.0])[None].repeat(f, 1).cuda()
    pitch = torch.linspace(-5, 5, f).cuda()
    pipe = AnimationPipeline(None, None, None, mv.unet, mv.pano_unet, mv, DDIMScheduler(**SCHEDULER_KWARGS))

    def run(clip_seed):
        cond = Conditioning(mk((2, 5, 32), 24), mk((2 * m, 5, 32), 25), mk((2, f, 4096, 8), clip_seed),
                            mk((2, 1, f, 4096, 8), clip_seed + 1).expand(-1, m, -1, -1, -1), rel, pitch, 8)
        torch.manual_seed(5)
        random.seed(6)
        a, b = pipe.denoise(pano, pers, pmask, vmask, pm, vm, cond, g["cams"], 50, 7.5, step_range=(0, 2))
        a, b = pipe.denoise(a, b, pmask, vmask, pm, vm, cond, g["cams"], 50, 7.5, step_range=(2, 4))     # resumes on the same graph
        return a, b

    monkeypatch.setenv("I360_CUDA_GRAPH", "0")
    e1, e2 = run(26), run(36)
    monkeypatch.setenv("I360_CUDA_GRAPH", "1")
    g1, g2 = run(26), run(36)               # second clip: same graph, conditioning reloaded, adapter graph replayed
    assert "_step_graph" in pipe.__dict__
    for (ea, eb), (ga, gb) in ((e1, g1), (e2, g2)):
        assert torch.equal(ea, ga) and torch.equal(eb, gb)
    assert not torch.equal(e1[0], e2[0])


def test_decode_video_streamed_equals_decode_video():
    """Output side: chunked decode with the D2H copies of chunk k under the decode of chunk k+1;
    fp32 video identical to decode_video(...).cpu(), uint8 frames identical to save_videos_grid's conversion, chunks
    delivered to the host callback in order."""
    import numpy as np
    from imagine360_b200.host.pipeline import AnimationPipeline
    from imagine360_b200.host.vae import AutoencoderKL
    g = load("vae.pt")
    vae = AutoencoderKL(**VAE_KW)
    load_native(vae, q(synth_state(g["shapes"], g["seed"]))[0])
    pipe = AnimationPipeline(vae, None, None, None, None, None, None)
    lat = qt(synth_tensor((1, 4, 10, 8, 16), 77, 0.2))[0]
    want = pipe.decode_video(lat).cpu()
    seen = []
    video, frames = pipe.decode_video_streamed(lat, frames_per_call=4, on_frames=lambda i, fr: seen.append((i, fr.clone())))
    assert video.shape == want.shape == (1, 3, 10, 64, 128) and torch.equal(video, want)
    ref_u8 = (want[0].permute(1, 2, 3, 0) * 255).numpy().astype(np.uint8)          # util.py:61-67 for one video
    assert np.array_equal(frames.numpy(), ref_u8)
    assert [i for i, _ in seen] == [0, 4, 8] and all(torch.equal(fr, frames[i:i + len(fr)]) for i, fr in seen)
