"""Calibrated GPU parity (SURVEY.md 8(c) tolerance policy).

BASELINE's "fp16 rtol 2e-3" is below ONE rounding of the production dtype (bf16: 2^-8 = 3.9e-3; the reference runs
bf16, inference_dual_p2e.py:378), so no 16-bit implementation -- the reference's own included -- can meet it after a
block of depth > 1.  What CAN be stated and checked:

 (1) three-way calibration, per block and for a whole dual-branch step at FULL channel widths: the oracle (reference
     algorithm) is run on the same GPU twice, in fp32 (the truth; TF32 off) and in bf16 through torch's library kernels
     (= what the reference's production path computes), and the native path must be as close to the truth as the
     reference's own bf16 path is:   err(native, fp32) <= 2 * err(bf16 oracle, fp32) + atol.   Both errors are printed.
 (2) kernels that accumulate in fp32 meet rtol 2e-3 BEFORE their single output rounding: element-wise
     |native - fp32| <= (2^-8 + 2e-3) |fp32| + atol  (2^-8 = the unit round-off of the single bf16 store).
 (3) the C5 extents (24 x 768 x 1536): 18 432-token self-attention, WarpAttn with 576 / 144 / 36 / 9-token views (the
     tile-padded bias path), temporal modules and the adapter at F = 24.
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_util import load, synth_state, synth_tensor, tiny_cameras  # noqa: E402
from test_blocks_gpu import BF, TINY, load_native, ncfhw, nhwc, q, qt  # noqa: E402
from test_host_modules import tiny_unet  # noqa: E402

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

K_BUDGET = 2.0      # native may be at most this many times as far from the fp32 truth as the reference's bf16 path
ATOL = 2e-3         # + this fraction of max|truth| (half a bf16 ulp of the largest element)


def _errs(a, ref):
    a, ref = a.float(), ref.float()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    assert torch.isfinite(a).all()
    d = a - ref
    return (d.abs().max() / (ref.abs().max() + 1e-12)).item(), (d.pow(2).mean().sqrt() / (ref.pow(2).mean().sqrt() + 1e-12)).item()


def calibrated(native, ref32, ref16, what, k=K_BUDGET, atol=ATOL):
    mn, rn = _errs(native, ref32)
    mb, rb = _errs(ref16, ref32)
    print(f"[calibrated] {what}: native max {mn:.3e} rms {rn:.3e} | bf16-oracle max {mb:.3e} rms {rb:.3e}")
    assert mn <= k * mb + atol, f"{what}: native max-err {mn:.3e} > {k} x bf16-oracle {mb:.3e} + {atol}"
    assert rn <= k * rb + atol / 4, f"{what}: native rms-err {rn:.3e} > {k} x bf16-oracle {rb:.3e} + {atol / 4}"


def pre_rounding_2e3(out, ref32, what, atol_frac, per_row=False):
    """|native - fp32| <= (2^-8 + 2e-3) |fp32| + atol_frac * rms(fp32): the bf16 store's round-off plus BASELINE's rtol.
    ``per_row``: the rms is taken over the last dim of each row (attention: the noise scales with the row's own output)."""
    out, ref = out.float(), ref32.float()
    assert out.shape == ref.shape
    err = (out - ref).abs()
    rms = ref.pow(2).mean(dim=-1, keepdim=True).sqrt() if per_row else ref.pow(2).mean().sqrt()
    tol = (2.0 ** -8 + 2e-3) * ref.abs() + atol_frac * rms
    bad = err > tol
    print(f"[pre-rounding] {what}: max err/tol {(err / tol).max().item():.3f}")
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} elements beyond rtol 2e-3 + one bf16 round-off (max excess {(err - tol).max().item():.3e})"


# ------------------------------------------------------------------------------------------------
# (1) blocks
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,halo,skip", [(320, 320, 2, 0), (640, 320, 0, 320), (1280, 1280, 2, 1280), (320, 640, 0, 0)])
def test_resnet_block_calibrated(cin, cout, halo, skip):
    from imagine360_b200.host import forward as Fw
    from imagine360_b200.host.unet3d import ResnetBlock3D
    from oracle import geometry as G, unet3d as OU
    from oracle.nn_ops import P
    m = ResnetBlock3D(in_channels=cin + skip, out_channels=cout, temb_channels=1280, groups=32, eps=1e-5)
    sd_n, sd_o = q(synth_state({k: list(v.shape) for k, v in m.state_dict().items()}, 1))
    load_native(m, sd_n)
    b, f, h, w = 2, 3, 8, 12
    x, xo = qt(synth_tensor((b, cin, f, h, w), 2))
    s, so = qt(synth_tensor((b, skip, f, h, w), 3)) if skip else (None, None)
    temb, tembo = qt(synth_tensor((b, 1280), 4))
    cfg = dict(groups=32, resnet_eps=1e-5)
    ref32 = G.unpad_pano(OU.resnet_block(G.pad_pano(torch.cat([xo, so], 1) if skip else xo, halo), tembo, P(sd_o), cfg), halo)
    ref16 = G.unpad_pano(OU.resnet_block(G.pad_pano(torch.cat([x, s], 1) if skip else x, halo), temb, P(sd_n), cfg), halo)
    tproj = (F.silu(temb).float() @ sd_o["time_emb_proj.weight"].t() + sd_o["time_emb_proj.bias"]).contiguous()
    out = Fw.resnet_block(nhwc(x), m, tproj, f, 32, skip=nhwc(s) if skip else None, halo=halo)
    calibrated(ncfhw(out, b), ref32, ref16, f"ResnetBlock3D {cin}+{skip}->{cout} halo {halo}")


@pytest.mark.parametrize("c,heads,hw", [(320, 5, (16, 16)), (640, 10, (8, 8)), (1280, 20, (4, 4))])
def test_spatial_transformer_calibrated(c, heads, hw):
    from imagine360_b200.host import forward as Fw
    from imagine360_b200.host.unet3d import Transformer3DModel
    from oracle import unet3d as OU
    from oracle.nn_ops import P
    dctx, n_ip = 1024, 64
    m = Transformer3DModel(heads, c // heads, c, dctx, 32, dctx, 1.0, n_ip)
    sd_n, sd_o = q(synth_state({k: list(v.shape) for k, v in m.state_dict().items()}, 5))
    load_native(m, sd_n)
    b, f = 2, 3
    x, xo = qt(synth_tensor((b, c, f, *hw), 6))
    ctx, ctxo = qt(synth_tensor((b, 77 + n_ip, dctx), 7))
    cfg = dict(groups=32, num_tokens=n_ip, ip_scale=1.0)
    ref32 = OU.spatial_transformer(xo, ctxo, P(sd_o), heads, cfg)
    ref16 = OU.spatial_transformer(x, ctx, P(sd_n), heads, cfg)
    out = Fw.spatial_transformer(nhwc(x), m, Fw.Context(ctx[:, :77], ctx[:, 77:]), f)
    calibrated(ncfhw(out, b), ref32, ref16, f"Transformer3DModel C={c}")


@pytest.mark.parametrize("c,frames", [(320, 16), (640, 16), (1280, 16), (320, 24), (1280, 24)])
def test_temporal_module_calibrated(c, frames):
    """F = 24 is the C5 clip length (temporal PE table length 64)."""
    from imagine360_b200.host import forward as Fw
    from imagine360_b200.host.unet3d import VanillaTemporalModule
    from oracle import unet3d as OU
    from oracle.nn_ops import P
    m = VanillaTemporalModule(c, num_attention_heads=8, num_transformer_block=1, temporal_position_encoding=True,
                              temporal_position_encoding_max_len=64)
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    sd_n, sd_o = q(synth_state(shapes, 8))
    load_native(m, sd_n)
    x, xo = qt(synth_tensor((2, c, frames, 3, 4), 9))
    for k in list(shapes):
        if k.endswith("pos_encoder.pe"):
            sd_o[k] = m.state_dict()[k].float()
            sd_n[k] = m.state_dict()[k]
    cfg = dict(groups=32, mm_heads=8, temporal_pe_max_len=64)
    ref32 = OU.temporal_module(xo, P(sd_o), cfg)
    ref16 = OU.temporal_module(x, P(sd_n), cfg)
    out = Fw.temporal_module(nhwc(x), m, frames)
    calibrated(ncfhw(out, 2), ref32, ref16, f"VanillaTemporalModule C={c} F={frames}")


# C3 levels: (320: 16x16 views, 32x64 pano) ...; C5 levels: 24x24 = 576, 12x12 = 144, 6x6 = 36, 3x3 = 9 tokens per view
@pytest.mark.parametrize("dim,m_,ph,eh,ew,anti", [(320, 20, 16, 32, 64, False), (640, 20, 8, 16, 32, True), (1280, 20, 4, 8, 16, False),
                                                  (320, 20, 24, 48, 96, True), (640, 20, 12, 24, 48, False),
                                                  (1280, 20, 6, 12, 24, True), (1280, 20, 3, 6, 12, False)])
def test_warp_attn_calibrated(dim, m_, ph, eh, ew, anti):
    from imagine360_b200.host.mvgen import WarpAttn
    from oracle import geometry as G
    from oracle import mvgen as OM
    from oracle.nn_ops import P
    w = WarpAttn(dim)
    shapes = {k: list(v.shape) for k, v in w.state_dict().items()}
    sd_n, sd_o = q(synth_state(shapes, 10))
    load_native(w, sd_n)
    cams = G.default_cameras()
    b, f = 1, 2
    pers, perso = qt(synth_tensor((b * m_, dim, f, ph, ph), 11))
    equi, equio = qt(synth_tensor((b, dim, f, eh, ew), 12))
    sd_o["pe.freq_bands"] = w.pe.freq_bands.float()
    sd_n["pe.freq_bands"] = w.pe.freq_bands
    p32, e32 = OM.warp_attn(perso, equio, cams, P(sd_o), anti, mask_dtype=None, grid_dtype=BF, pe_dtype=BF)
    p16, e16 = OM.warp_attn(pers, equi, cams, P(sd_n), anti)        # everything in bf16, as production
    pn, en = w.forward_native(nhwc(pers), nhwc(equi), cams, b, m_, f, anti)
    calibrated(ncfhw(pn, b * m_), p32, p16, f"WarpAttn dim {dim} views {ph}x{ph} pers<-equi")
    calibrated(ncfhw(en, b), e32, e16, f"WarpAttn dim {dim} pano {eh}x{ew} equi<-pers")


@pytest.mark.parametrize("frames", [16, 24])
def test_adapter_calibrated(frames):
    """F = 24: two avg_pool1d(4) stages 24 -> 6 -> 1 (resampler.py:251,:264)."""
    from imagine360_b200.host import mvgen as M
    from oracle import unet3d as OU
    from oracle.nn_ops import P
    u = tiny_unet()
    shapes = {k: list(v.shape) for k, v in u.state_dict().items()}
    sd_n, sd_o = q(synth_state(shapes, 13))
    load_native(u, sd_n)
    feats, featso = qt(synth_tensor((2, frames, 4096, 8), 14))
    cfg = dict(tproj_heads=8, adapter_heads=12, adapter_dim_head=64)

    def run(ft, sd):
        y1 = OU.temporal_projection(ft, P(sd, "temporal_proj."), cfg)
        b, f, n, d = y1.shape
        assert f == 1
        return OU.resampler(y1.reshape(b, f * n, d), P(sd, "image_proj_model."), cfg)

    calibrated(M.ip_tokens_clean(u, feats), run(featso, sd_o), run(feats, sd_n), f"adapter F={frames}")


# ------------------------------------------------------------------------------------------------
# (1b) one dual-branch step, FULL channel widths (320/640/1280, heads 5/10/20/20, 64 IP tokens, 20 views), reduced extent
# ------------------------------------------------------------------------------------------------
def test_full_width_dual_step_calibrated():
    import random
    from imagine360_b200.host.config import FULL_UNET_KWARGS
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    from imagine360_b200.host.pipeline import random_init_, synthetic_inputs
    from imagine360_b200.host.unet3d import UNet3DConditionModel
    from oracle import mvgen as OM
    torch.manual_seed(0)
    with torch.device("cuda"):
        mv = MultiViewBaseModel(UNet3DConditionModel(**FULL_UNET_KWARGS), UNet3DConditionModel(**FULL_UNET_KWARGS)).to(BF)
    random_init_(mv)
    frames, m = 16, 20
    inp = synthetic_inputs(frames=frames, pano_hw=(256, 512), views=m, device="cuda", seed=5)     # the C2 extent
    cond = inp["cond"]
    xin_pano = torch.cat([inp["pano_latent"], inp["pano_mask"], inp["pano_masked"]], 1)
    xin_pers = torch.cat([inp["pers_latent"], inp["pers_masks"], inp["pers_masked"]], 2)
    lat, plat = torch.cat([xin_pers] * 2), torch.cat([xin_pano] * 2)
    t = torch.tensor([481], device="cuda")
    fps_pano = torch.tensor([8, 8], device="cuda")
    fps_pers = fps_pano[:, None].repeat(1, m)
    rel, pitch = cond.rel_pos[None].repeat(2, 1, 1), cond.pitch[None].repeat(2, 1)
    draws = [False, True, False, False, True, False, True]
    g = torch.Generator(device="cuda").manual_seed(9)
    n_pano = torch.randn(2, 64, 1024, device="cuda", generator=g).to(BF)
    n_pers = torch.randn(2 * m, 64, 1024, device="cuda", generator=g).to(BF)
    cams = {k: [float(x) for x in inp["cameras"][k].reshape(-1)] for k in ("FoV", "theta", "phi")}
    ns, npn = mv(latents=lat, pano_latent=plat, timestep=t, prompt_embd=cond.text_pers, pano_prompt_embd=cond.text_pano,
                 cameras=inp["cameras"], use_fps_condition=True, use_ip_plus_cross_attention=True, fps_tensor_pano=fps_pano,
                 fps_tensor_pers=fps_pers, reference_images_clip_feat_pano=cond.feats_pano,
                 reference_images_clip_feat_pers=cond.feats_pers, relative_position_tensor=rel, pitchs_tensor=pitch,
                 antipodal_draws=draws, ip_noise=(n_pano, n_pers))
    sd_b = dict(mv.state_dict())
    ys16, yp16 = OM.mv_forward(sd_b, lat, plat, t, cond.text_pers, cond.text_pano, cams, fps_pano, fps_pers, cond.feats_pano,
                               cond.feats_pers, rel.to(BF), pitch.to(BF), draws, n_pano, n_pers)
    sd_f = {k: v.float() for k, v in sd_b.items()}
    f32 = lambda x: x.float()   # noqa: E731
    ys32, yp32 = OM.mv_forward(sd_f, f32(lat), f32(plat), t, f32(cond.text_pers), f32(cond.text_pano), cams, fps_pano, fps_pers,
                               f32(cond.feats_pano), f32(cond.feats_pers), rel, pitch, draws, f32(n_pano), f32(n_pers),
                               grid_dtype=BF, pe_dtype=BF)
    calibrated(ns, ys32, ys16, "full-width dual step, pers output")
    calibrated(npn, yp32, yp16, "full-width dual step, pano output")


def test_tiny_dual_step_calibrated():
    """the committed golden configuration (tests/golden/mvgen.pt), three-way"""
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    from oracle import mvgen as OM
    from test_oracle_golden import mvgen_inputs
    g = load("mvgen.pt")
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet())
    sd_n, sd_o = q(synth_state(g["shapes"], g["seed"]))
    load_native(mv, sd_n)
    inp = {k: qt(v) for k, v in mvgen_inputs(g).items()}
    nat, ora = {k: v[0] for k, v in inp.items()}, {k: v[1] for k, v in inp.items()}
    for k in g["shapes"]:
        if k.endswith(("pos_encoder.pe", "pe.freq_bands")):
            sd_o[k] = mv.state_dict()[k].float()
            sd_n[k] = mv.state_dict()[k]
    t = torch.tensor([g["t"]]).cuda()
    fps_pano, fps_pers = torch.tensor([8, 8]).cuda(), torch.tensor([[8, 8], [8, 8]]).cuda()

    def run(sd, d, **kw):
        return OM.mv_forward(sd, d["latents"], d["pano_latent"], t, d["prompt_embd"], d["pano_prompt_embd"], g["cams"], fps_pano,
                             fps_pers, d["feats_pano"], d["feats_pers"], d["rel_pos"], d["pitch"], g["draws"], d["ip_noise_pano"],
                             d["ip_noise_pers"], cfg=TINY, **kw)

    ys32, yp32 = run(sd_o, ora, grid_dtype=BF, pe_dtype=BF)
    ys16, yp16 = run(sd_n, nat)
    ns, np_ = mv(latents=nat["latents"], pano_latent=nat["pano_latent"], timestep=t, prompt_embd=nat["prompt_embd"],
                 pano_prompt_embd=nat["pano_prompt_embd"], cameras=g["cams"], use_fps_condition=True,
                 use_ip_plus_cross_attention=True, fps_tensor_pano=fps_pano, fps_tensor_pers=fps_pers,
                 reference_images_clip_feat_pano=nat["feats_pano"], reference_images_clip_feat_pers=nat["feats_pers"],
                 relative_position_tensor=nat["rel_pos"], pitchs_tensor=nat["pitch"], antipodal_draws=g["draws"],
                 ip_noise=(nat["ip_noise_pano"], nat["ip_noise_pers"]))
    calibrated(ns, ys32, ys16, "tiny dual step, pers output")
    calibrated(np_, yp32, yp16, "tiny dual step, pano output")


# ------------------------------------------------------------------------------------------------
# (2) fp32-accumulating kernels: rtol 2e-3 before the output rounding
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1000, 320, 320), (4096, 1280, 1280), (513, 2560, 1280), (640, 1920, 640), (130, 1024, 4096)])
def test_gemm_meets_2e3_before_rounding(M, N, K):
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g).bfloat16()
    pre_rounding_2e3(ops.gemm(a, w, bias=bias), a.double() @ w.double().t() + bias.double(), f"gemm {M}x{N}x{K}", 1e-4)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(3, 32, 32, 320, 320), (2, 64, 132, 320, 320), (2, 8, 20, 2560, 1280)])
def test_conv_meets_2e3_before_rounding(B, H, W, Cin, Cout):
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + H + Cin)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g).bfloat16()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    pre_rounding_2e3(ops.conv3x3(x, ops.pack_conv3x3(w), bias=bias), ref, f"conv {Cin}->{Cout}", 1e-4)


def _sdpa32(qh, kh, vh, bias=None):
    return F.scaled_dot_product_attention(qh.float(), kh.float(), vh.float(), attn_mask=None if bias is None else bias.float())


@pytest.mark.parametrize("imgs,N,heads", [(2, 1024, 5), (1, 8192, 2), (1, 18432, 5)])
def test_self_attention_meets_2e3_before_rounding(imgs, N, heads):
    """N = 18 432 is the C5 panorama level-0 sequence (96 x 192 latent).  P is a bf16 MMA operand (as in every fused
    attention kernel, xformers / SDPA included): each p_ij carries an independent rounding error of up to 2^-8 relative
    (rms ~1.6e-3), so an output element is off by ~1.6e-3 of its row's output rms (1 sigma) whatever the number of keys;
    six sigma over 10^6..10^7 elements is the absolute term 1e-2 * rms(row) (1.5e-2 for the hd-32 biased kernel) (measured on the B200 at 7e-3: 14 of 10^6
    elements outside, worst 1.13 x).  The relative part stays 2e-3 + one round-off."""
    from imagine360_b200 import ops
    hd = 64
    C = heads * hd
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn(imgs * N, 3 * C, device="cuda", generator=g).bfloat16()
    out = torch.zeros(imgs * N, C, device="cuda", dtype=BF)
    ops.attention(ops.seq_view(qkv, imgs, N, 0), ops.seq_view(qkv, imgs, N, C), ops.seq_view(qkv, imgs, N, 2 * C),
                  ops.seq_view(out, imgs, N), heads, hd, imgs)
    qh, kh, vh = (qkv[:, i * C:(i + 1) * C].reshape(imgs, N, heads, hd).transpose(1, 2) for i in range(3))
    ref = _sdpa32(qh, kh, vh).transpose(1, 2).reshape(imgs * N, C)
    pre_rounding_2e3(out.reshape(imgs * N, heads, hd), ref.reshape(imgs * N, heads, hd), f"self-attention N={N}", 1e-2, per_row=True)


# ------------------------------------------------------------------------------------------------
# (3) C5 extents of the biased (WarpAttn) attention kernel: views of 576 / 144 / 36 / 9 tokens
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,ph,eh,ew,heads", [(20, 24, 48, 96, 10), (20, 12, 24, 48, 20), (20, 6, 12, 24, 40), (20, 3, 6, 12, 40)])
def test_warp_attention_c5_views(m, ph, eh, ew, heads):
    from imagine360_b200 import ops
    hd, b, Fr = 32, 1, 2
    C = heads * hd
    hw, EN = ph * ph, eh * ew
    g = torch.Generator(device="cuda").manual_seed(ph)

    def rnd(*s):
        return torch.randn(*s, device="cuda", generator=g).bfloat16()

    pers, pers_kv = rnd(b * m * Fr * hw, C), rnd(b * m * Fr * hw, 2 * C)
    equi, equi_kv = rnd(b * Fr * EN, C), rnd(b * Fr * EN, 2 * C)
    bias_e, bias_p = rnd(EN, m * hw), rnd(m * hw, EN)
    out_e = torch.full_like(equi, float("nan"))
    out_p = torch.full_like(pers, float("nan"))
    ops.attention(ops.seq_view(equi, b * Fr, EN), ops.multiview_view(pers_kv, b, m, Fr, hw, 0),
                  ops.multiview_view(pers_kv, b, m, Fr, hw, C), ops.seq_view(out_e, b * Fr, EN), heads, hd, b * Fr, bias=bias_e)
    ops.attention(ops.multiview_view(pers, b, m, Fr, hw), ops.seq_view(equi_kv, b * Fr, EN, 0), ops.seq_view(equi_kv, b * Fr, EN, C),
                  ops.multiview_view(out_p, b, m, Fr, hw), heads, hd, b * Fr, bias=bias_p)

    def to_bf(t, c):   # (b m f t) c -> (b f) (m t) c
        return t.reshape(b, m, Fr, hw, c).permute(0, 2, 1, 3, 4).reshape(b * Fr, m * hw, c)

    def heads_of(t):
        return t.reshape(t.shape[0], t.shape[1], heads, hd).transpose(1, 2)

    pkv, ekv = to_bf(pers_kv, 2 * C), equi_kv.reshape(b * Fr, EN, 2 * C)
    ref_e = _sdpa32(heads_of(equi.reshape(b * Fr, EN, C)), heads_of(pkv[..., :C]), heads_of(pkv[..., C:]), bias_e)
    ref_p = _sdpa32(heads_of(to_bf(pers, C)), heads_of(ekv[..., :C]), heads_of(ekv[..., C:]), bias_p)
    pre_rounding_2e3(out_e.reshape(b * Fr, EN, heads, hd).transpose(1, 2), ref_e, f"warp equi<-pers views {hw}", 1.5e-2, per_row=True)
    pre_rounding_2e3(to_bf(out_p, C).reshape(b * Fr, m * hw, heads, hd).transpose(1, 2), ref_p, f"warp pers<-equi views {hw}", 1.5e-2, per_row=True)
