"""Host wiring of ``forward.resnet_block`` on the CPU: the op sequence the host issues (GroupNorm over the concat / the padded
panorama tensor, conv1 + temb, norm2, conv2 + fused shortcut / residual + crop, 1 / output_scale_factor) is run with the
torch restatements of ``imagine360_b200/debug.py`` in place of the CUDA kernels and compared with the oracle's ResnetBlock3D
(animatediff/models/resnet.py:221-254 restated in oracle/unet3d.py).  The debug ops stand in for the kernels HERE only so
that the wiring can be checked without a GPU; the kernels themselves are checked in the ``-m gpu`` tests.  Both statistics
routes are covered: the statistics pass (default) and the opt-in conv -> GroupNorm route (I360_CONV_GN_STATS=1), where norm2
reads conv1's per-channel sums and conv2 leaves the sums of the block's result for the Transformer3DModel norm that follows."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_util import synth_state, synth_tensor  # noqa: E402

torch.set_grad_enabled(False)
BF = torch.bfloat16


def _nhwc(x5):     # [b,c,f,h,w] -> [(b f),h,w,c]
    b, c, f, h, w = x5.shape
    return x5.permute(0, 2, 3, 4, 1).reshape(b * f, h, w, c).contiguous()


def _ncfhw(x4, b):
    n, h, w, c = x4.shape
    return x4.reshape(b, n // b, h, w, c).permute(0, 4, 1, 2, 3)


@pytest.mark.parametrize("conv_gn", [False, True])
@pytest.mark.parametrize("cin,cout,halo,skip", [(64, 64, 0, 0), (64, 128, 0, 0), (128, 64, 2, 64), (320, 320, 2, 0)])
def test_resnet_block_wiring(cin, cout, halo, skip, conv_gn, monkeypatch):
    from imagine360_b200 import debug
    from imagine360_b200.host import forward as Fw
    from imagine360_b200.host.unet3d import ResnetBlock3D
    from oracle import geometry as G, unet3d as OU
    from oracle.nn_ops import P
    monkeypatch.setattr(Fw, "CONV_GN", conv_gn)
    m = ResnetBlock3D(in_channels=cin + skip, out_channels=cout, temb_channels=96, groups=32, eps=1e-5)
    sd = {k: v.to(BF) for k, v in synth_state({k: list(v.shape) for k, v in m.state_dict().items()}, 1).items()}
    m.load_state_dict(sd)
    m = m.to(BF)
    sd_o = {k: v.float() for k, v in sd.items()}
    b, f, h, w = 2, 3, 8, 12
    x = synth_tensor((b, cin, f, h, w), 2).to(BF)
    s = synth_tensor((b, skip, f, h, w), 3).to(BF) if skip else None
    temb = synth_tensor((b, 96), 4).to(BF)
    xin = torch.cat([x, s], 1).float() if skip else x.float()
    ref = G.unpad_pano(OU.resnet_block(G.pad_pano(xin, halo), temb.float(), P(sd_o), dict(groups=32, resnet_eps=1e-5)), halo)
    tproj = (F.silu(temb).float() @ sd_o["time_emb_proj.weight"].t() + sd_o["time_emb_proj.bias"]).contiguous()
    with debug.reference_ops("conv3x3", "groupnorm"):
        out = Fw.resnet_block(_nhwc(x), m, tproj, f, 32, skip=_nhwc(s) if skip else None, halo=halo, out_stats=True)
    got = _ncfhw(out, b).float()
    assert got.shape == ref.shape
    # bf16 storage between the five ops, fp32 inside them: a few bf16 round-offs
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 2e-2
    st = getattr(out, "_i360_chan_stats", None)
    assert (st is not None) == conv_gn
    if conv_gn:      # the sums conv2 leaves are those of the stored (cropped, scaled) result, per image and channel
        v = out.double().view(b * f, -1, cout)
        assert torch.allclose(st, torch.stack([v.sum(1), (v * v).sum(1)], -1))


def test_conv_gn_route_needs_32_pixels(monkeypatch):
    """A warp's 32 tile rows must lie in one image: below 32 pixels per image the host keeps the statistics pass."""
    from imagine360_b200.host import forward as Fw
    monkeypatch.setattr(Fw, "CONV_GN", True)
    assert Fw._chan_stats_ok(4, 8) and Fw._chan_stats_ok(64, 132) and not Fw._chan_stats_ok(4, 4) and not Fw._chan_stats_ok(2, 8)
    monkeypatch.setattr(Fw, "CONV_GN", False)
    assert not Fw._chan_stats_ok(64, 64)


def _state(m, seed):
    sd = {k: v.to(BF) for k, v in synth_state({k: list(v.shape) for k, v in m.state_dict().items()}, seed).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(("pos_encoder.pe", "pe.freq_bands")) for k in missing)
    return m.to(BF), {k: v.float() for k, v in sd.items()}


@pytest.mark.parametrize("c,heads,n_ip,hw,ln_fold", [(128, 2, 16, (4, 6), True), (128, 2, 16, (4, 6), False), (64, 2, 4, (4, 4), True),
                                                    (320, 5, 64, (4, 4), True)])
def test_spatial_transformer_wiring(c, heads, n_ip, hw, ln_fold, monkeypatch):
    """Transformer3DModel / BasicTransformerBlock (animatediff/models/attention.py:246-301,:461-508): GroupNorm -> proj_in ->
    self-attention on the fused QKV projection -> text + image-prompt cross-attention (head_dim 64: the fused kernel's route;
    head_dim 32: two attention calls with accumulate) -> GEGLU feed-forward -> proj_out + residual, with the LayerNorms folded
    into the projections or run on their own."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_ops import cpu_ops
    from imagine360_b200.host import forward as Fw
    from imagine360_b200.host.unet3d import Transformer3DModel
    from oracle import unet3d as OU
    from oracle.nn_ops import P
    monkeypatch.setattr(Fw, "LN_FOLD", ln_fold)
    dctx = 128
    m, sd_o = _state(Transformer3DModel(heads, c // heads, c, dctx, 32, dctx, 1.0, n_ip), 5)
    b, f = 2, 3
    x = synth_tensor((b, c, f, *hw), 6).to(BF)
    ctx = synth_tensor((b, 7 + n_ip, dctx), 7).to(BF)
    ref = OU.spatial_transformer(x.float(), ctx.float(), P(sd_o), heads, dict(groups=32, num_tokens=n_ip, ip_scale=1.0))
    calls = {}
    with cpu_ops(calls):
        out = Fw.spatial_transformer(_nhwc(x), m, Fw.Context(ctx[:, :7], ctx[:, 7:]), f)
    assert (calls.get("cross_attention_text_ip", 0) > 0) == (c // heads == 64)
    # folded: the projections that pay take the statistics from their producer (the rest keep a LayerNorm pass); not folded: the
    # block's three LayerNorms run on their own
    assert (calls.get("gemm_ln", 0) > 0) == ln_fold and (ln_fold or calls["layernorm"] == 3)
    got = _ncfhw(out, b).float()
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 2.5e-2       # the GPU test's budget: ~8 GEMMs deep


@pytest.mark.parametrize("c,heads,frames", [(64, 2, 5), (320, 8, 16), (64, 2, 24)])
def test_temporal_module_wiring(c, heads, frames):
    """VanillaTemporalModule (animatediff/models/motion_module.py:52-429): GroupNorm -> proj_in -> two temporal self-attentions with
    the sinusoidal PE added to the normalised tokens -> feed-forward -> proj_out + residual, sequences over the frame axis read
    in place from the token-major projection output."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_ops import cpu_ops
    from imagine360_b200.host import forward as Fw
    from imagine360_b200.host.unet3d import VanillaTemporalModule
    from oracle import unet3d as OU
    from oracle.nn_ops import P
    m = VanillaTemporalModule(c, num_attention_heads=heads, num_transformer_block=1, temporal_position_encoding=True,
                              temporal_position_encoding_max_len=64)
    m, sd_o = _state(m, 8)                       # re-randomises the zero-initialised proj_out
    for k, v in m.state_dict().items():          # the reference adds the bf16-cast PE buffer; give the oracle the same values
        if k.endswith("pos_encoder.pe"):
            sd_o[k] = v.float()
    b = 2
    x = synth_tensor((b, c, frames, 3, 4), 9).to(BF)
    ref = OU.temporal_module(x.float(), P(sd_o), dict(groups=32, mm_heads=heads, temporal_pe_max_len=64))
    calls = {}
    with cpu_ops(calls):
        out = Fw.temporal_module(_nhwc(x), m, frames)
    assert calls["temporal_attention"] == 2
    got = _ncfhw(out, b).float()
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 2.5e-2


@pytest.mark.parametrize("dim,m_,ph,eh,ew,anti", [(64, 3, 4, 8, 16, False), (64, 3, 4, 8, 16, True), (128, 2, 8, 16, 32, False)])
def test_warp_attn_wiring(dim, m_, ph, eh, ew, anti):
    """WarpAttn (src/modules/attn_perspano.py:22-99): soft visibility masks from get_merged_masks (+ the antipodal variant),
    spherical position encodings sampled through e2p / p2e grids, the two cross-attentions (views -> panorama, panorama ->
    views) as biased attention over (view, token) sequences, residual + feed-forward per side."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_ops import cpu_ops
    from golden_util import tiny_cameras
    from imagine360_b200.host.mvgen import WarpAttn
    from oracle import mvgen as OM
    from oracle.nn_ops import P
    w, sd_o = _state(WarpAttn(dim), 10)
    cams = tiny_cameras(m_)
    b, f = 2, 2
    pers = synth_tensor((b * m_, dim, f, ph, ph), 11).to(BF)
    equi = synth_tensor((b, dim, f, eh, ew), 12).to(BF)
    sd_o["pe.freq_bands"] = w.pe.freq_bands.float()      # bf16-cast buffer, as after model.to(bfloat16)
    po, eo = OM.warp_attn(pers.float(), equi.float(), cams, P(sd_o), anti, mask_dtype=None, grid_dtype=BF, pe_dtype=BF)
    calls = {}
    with cpu_ops(calls):
        pn, en = w.forward_native(_nhwc(pers), _nhwc(equi), cams, b, m_, f, anti)
    assert calls["attention"] == 2 and calls["grid_sample"] > 0
    for got, ref, n in ((pn, po, b * m_), (en, eo, b)):
        got = _ncfhw(got, n).float()
        assert ((got - ref).abs().max() / ref.abs().max()).item() < 2.5e-2


def test_single_unet_forward_vs_oracle_and_reference():
    """``UNet3DConditionModel.forward`` (animatediff/models/unet.py:632, BASELINE.json configs[1]'s graph) at the tiny widths of
    tests/golden/unet3d.pt: the host wiring with emulated kernels against the oracle AND against the output the UNMODIFIED
    reference produced for the same weights and inputs (tools/make_golden.py)."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_ops import cpu_ops
    from golden_util import load
    from test_host_modules import tiny_unet
    from test_oracle_golden import TINY
    from oracle import unet3d as OU
    g = load("unet3d.pt")
    u, sd_o = _state(tiny_unet(), g["seed"])
    x = synth_tensor((1, 9, 4, 8, 16), g["x_seed"]).to(BF)
    ctx = synth_tensor((1, 21, 32), g["ctx_seed"]).to(BF)
    for k, v in u.state_dict().items():
        if k.endswith("pos_encoder.pe"):
            sd_o[k] = v.float()
    t, fps = torch.tensor([g["t"]]), torch.tensor([g["fps"]])
    ref = OU.unet3d_forward(sd_o, x.float(), t, ctx.float(), cfg=TINY, fps=fps)
    with cpu_ops():
        out = u(x, t, ctx, use_fps_condition=True, fps_tensor=fps).sample
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
    assert rel(out, ref) < 5e-2                    # ~60 ops deep with bf16 storage between them
    assert rel(out, g["y"]) < 6e-2                 # the unmodified reference's fp32 output


def test_geometry_cache_is_bounded_and_keeps_what_a_graph_holds():
    """The mask / PE / grid tables are cached per camera set; ``prune_cache`` keeps the most recently used sets and
    ``cached_entries`` is what a captured step graph holds on to (it reads the tables by address)."""
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_ops import cpu_ops
    from imagine360_b200.host import geometry as G
    G._grid_cache.clear()
    G._set_last_use.clear()
    fb = (2.0 ** torch.arange(4)).to(BF)
    sets = [dict(FoV=[90, 90], theta=[10.0 * i, 100.0 + i], phi=[5.0, -20.0]) for i in range(6)]
    with cpu_ops():
        for c in sets:
            G.warp_biases(4, 4, 8, 16, c, "cpu", False)
            G.spherical_pe_tables(fb, 4, 4, 8, 16, c, "cpu")
        held = G.cached_entries(sets[0])
        assert len(held) >= 4                                  # e2p grid, p2e grid + mask, bias pair, PE pair
        G.warp_biases(4, 4, 8, 16, sets[1], "cpu", False)      # set 1 becomes the most recently used
        G.prune_cache(max_sets=3)
        alive = {k[1] for k in G._grid_cache}
        assert alive == {G.camera_lists(c) for c in (sets[1], sets[4], sets[5])}
        assert G.cached_entries(sets[0]) == [] and all(isinstance(v, (tuple, torch.Tensor)) for v in held)      # still referenced
        again = G.warp_biases(4, 4, 8, 16, sets[0], "cpu", False)                                              # rebuilt on demand
        assert torch.equal(again[0], next(v for v in held if isinstance(v, tuple) and v[0].shape == again[0].shape)[0])
    G._grid_cache.clear()
    G._set_last_use.clear()
