"""GPU parity of the tcgen05 GEMM / implicit-GEMM conv engine against fp32 torch math.

These are floating-point kernels, so the checker is a plain fp32 torch evaluation of the same op on
the same bf16 inputs (the CPU oracle modules are compared at block level in test_blocks_gpu.py).
Tolerance: bf16 output rounding (2^-8 relative) on top of fp32 accumulation-order noise.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(out, ref, what, rtol=1.0 / 128, atol_scale=2e-3):
    out = out.float()
    ref = ref.float()
    atol = atol_scale * ref.abs().max().item() + 1e-6
    err = (out - ref).abs()
    bad = err > (atol + rtol * ref.abs())
    assert not bad.any(), f"{what}: {bad.sum().item()} / {bad.numel()} mismatches, max err {err.max().item():.4g}, ref max {ref.abs().max().item():.4g}"


@pytest.mark.parametrize("M,N,K", [
    (128, 64, 64), (256, 128, 128), (300, 320, 320), (1000, 640, 320), (4096, 1280, 1280),
    (77, 320, 1024), (2, 1280, 320), (5000, 960, 320), (513, 2560, 1280), (640, 1920, 640),
    (130, 1024, 4096), (128, 72, 200),
])
def test_gemm_plain(M, N, K):
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    out = ops.gemm(a, w)
    _close(out, a.float() @ w.float().t(), f"gemm {M}x{N}x{K}")


def test_gemm_epilogues():
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 1500, 640, 320
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g).bfloat16()
    resid = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    rowvec = torch.randn(M // 100, N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias.float()
    _close(ops.gemm(a, w, bias=bias), ref, "bias")
    _close(ops.gemm(a, w, bias=bias, resid=resid), ref + resid.float(), "bias+resid")
    _close(ops.gemm(a, w, bias=bias, act=ops.ACT_GELU), F.gelu(ref), "gelu")
    _close(ops.gemm(a, w, bias=bias, act=ops.ACT_SILU), F.silu(ref), "silu")
    rv = rowvec.repeat_interleave(100, dim=0)
    _close(ops.gemm(a, w, bias=bias, rowvec=rowvec, rowvec_div=100, out_scale=0.5), (ref + rv) * 0.5, "rowvec+scale")
    # strided A view and strided output view
    big = torch.randn(M, 2 * K, device="cuda", generator=g).bfloat16()
    outbig = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(big[:, K:], w, out=outbig[:, N:])
    _close(outbig[:, N:], big[:, K:].float() @ w.float().t(), "strided views")
    assert outbig[:, :N].abs().max().item() == 0


@pytest.mark.parametrize("C", [320, 640, 1280])
def test_gemm_geglu(C):
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(C)
    M = 777
    a = torch.randn(M, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(8 * C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    b = torch.randn(8 * C, device="cuda", generator=g).bfloat16()
    wp, bp = ops.pack_geglu(w, b)
    out = ops.gemm(a, wp, bias=bp, act=ops.ACT_GEGLU)
    h = a.float() @ w.float().t() + b.float()
    val, gate = h.chunk(2, dim=-1)
    _close(out, val * F.gelu(gate), f"geglu {C}")


def test_geglu_gate_function_accuracy():
    """The GEGLU epilogue's gelu (A-S 7.1.28 erf by default; the sigmoid-quintic experiment gelu_sig2 meets the same bound)
    against erf-GELU on gates sweeping [-30, 30], isolated with an identity weight (accumulators == inputs):
    |error| <= 3e-5 + one bf16 rounding of the result."""
    from imagine360_b200 import ops
    M = 4096
    gates = torch.linspace(-30, 30, M * 128, device="cuda").view(M, 128).bfloat16()
    vals = torch.ones(M, 128, device="cuda").bfloat16()
    a = torch.cat([vals, gates], 1).contiguous()
    w = torch.eye(256, device="cuda").bfloat16()
    wp, _ = ops.pack_geglu(w, None)
    out = ops.gemm(a, wp, act=ops.ACT_GEGLU).float()
    ref = F.gelu(gates.double()).float()
    err = (out - ref).abs()
    bound = 3e-5 + 2.0 ** -8 * ref.abs()
    assert (err <= bound).all(), (err - bound).max().item()
    # the approximation itself (before the bf16 rounding of the output) stays far below that rounding wherever |gelu| > 0.02
    assert err[ref.abs() > 0.02].max().item() <= (2.0 ** -8 * ref.abs()[ref.abs() > 0.02]).max().item()


@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (2, 16, 16, 64, 64), (3, 32, 32, 320, 320), (5, 8, 8, 640, 1280), (9, 4, 4, 1280, 1280),
    (2, 64, 132, 320, 320), (2, 32, 68, 960, 640), (2, 8, 20, 2560, 1280), (1, 24, 40, 128, 256),
    (2, 16, 36, 64, 8),
])
def test_conv3x3(B, H, W, Cin, Cout):
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 131 + H * 17 + Cin + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g).bfloat16()
    out = ops.conv3x3(x, ops.pack_conv3x3(w), bias=bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1).permute(0, 2, 3, 1)
    _close(out, ref, f"conv {B}x{H}x{W} {Cin}->{Cout}")


def test_conv3x3_fused_shortcut_temb_crop():
    """conv2 of a pano ResnetBlock3D: 3x3 on the padded domain + 1x1 shortcut over two concatenated
    sources + temb + crop of the circular halo (MVGenModel.py:276-281, resnet.py:246-251)."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, Cin, Cout, C2, C3, crop, F_ = 4, 16, 36, 128, 192, 64, 128, 2, 2
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    x2 = torch.randn(B, H, W - 2 * crop, C2, device="cuda", generator=g).bfloat16()
    x3 = torch.randn(B, H, W - 2 * crop, C3, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    ws = (torch.randn(Cout, C2 + C3, 1, 1, device="cuda", generator=g) / (C2 + C3) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g).bfloat16()
    temb = torch.randn(B // F_, Cout, device="cuda", generator=g)
    resid = torch.randn(B, H, W - 2 * crop, Cout, device="cuda", generator=g).bfloat16()
    out = ops.conv3x3(x, ops.pack_conv3x3(w, ws), bias=bias, x2=x2, x3=x3, resid=resid, rowvec=temb,
                      rowvec_div=F_, crop=crop)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1)[..., crop:-crop]
    ref = ref + F.conv2d(torch.cat([x2, x3], -1).float().permute(0, 3, 1, 2), ws.float())
    ref = ref + temb.repeat_interleave(F_, 0)[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + resid.float()
    _close(out, ref, "fused conv")


@pytest.mark.parametrize("B,H,W,Cin,Cout,mode", [
    (2, 32, 32, 320, 320, "bias"), (2, 32, 32, 320, 320, "resid"), (4, 32, 32, 320, 320, "rowvec"), (1, 64, 136, 128, 128, "bias"),
    (2, 32, 36, 192, 160, "resid"), (2, 16, 16, 640, 640, "bias"), (1, 64, 136, 128, 8, "bias"), (2, 64, 136, 320, 320, "crop"),
    (2, 32, 72, 640, 320, "shortcut_crop"), (3, 48, 40, 72, 256, "bias"), (1, 128, 272, 256, 256, "resid"),
])
def test_conv3x3_halo_tiles(B, H, W, Cin, Cout, mode):
    """The halo-tile variant (one 18 x 10 TMA box per 64-channel block shared by the nine taps, shifted UMMA descriptors):
    every epilogue it is instantiated for, W / H tails, a partial last channel block, the cropped pano output."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 7 + H + W + Cin + Cout)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x = rn(B, H, W, Cin).bfloat16()
    w = (rn(Cout, Cin, 3, 3) / (9 * Cin) ** 0.5).bfloat16()
    bias = rn(Cout).bfloat16()
    crop = 2 if "crop" in mode else 0
    kw, ref = {}, F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), padding=1)
    if crop:
        ref = ref[..., crop:-crop]
    wp = ops.pack_conv3x3(w)
    if mode in ("resid", "crop"):
        kw["resid"] = rn(B, H, W - 2 * crop, Cout).bfloat16()
        ref = ref + kw["resid"].float().permute(0, 3, 1, 2)
    if mode == "rowvec":
        kw["rowvec"], kw["rowvec_div"] = rn(B // 2, Cout), 2
        ref = ref + kw["rowvec"].repeat_interleave(2, 0)[:, :, None, None]
    if mode == "shortcut_crop":
        x2, x3 = rn(B, H, W - 2 * crop, 64).bfloat16(), rn(B, H, W - 2 * crop, 96).bfloat16()
        ws = (rn(Cout, 160, 1, 1) / 160 ** 0.5).bfloat16()
        wp = ops.pack_conv3x3(w, ws)
        kw.update(x2=x2, x3=x3)
        ref = ref + F.conv2d(torch.cat([x2, x3], -1).float().permute(0, 3, 1, 2), ws.float())
    # W / H tails and the fused 1x1 sources are outside the production selection rule: widen it for this test so that
    # those kernel paths stay covered
    ops.conv3x3_halo_policy(1, 1.5, 1, 0)
    try:
        assert ops.conv3x3_uses_halo(B, H, W, Cin, "resid" in kw, "rowvec" in kw, "x2" in kw), "meant to take the halo kernels"
        out = ops.conv3x3(x, wp, bias=bias, crop=crop, **kw)
        out2 = ops.conv3x3(x, wp, bias=bias, crop=crop, **kw)
    finally:
        ops.conv3x3_halo_policy(1, 1.04, 0, 64)
    _close(out, ref.permute(0, 2, 3, 1), f"halo conv {mode} {B}x{H}x{W} {Cin}->{Cout}")
    assert torch.equal(out, out2)
    # the tap-by-tap kernels on the same problem agree to accumulation-order noise
    ops.conv3x3_halo_policy(0)
    try:
        out3 = ops.conv3x3(x, wp, bias=bias, crop=crop, **kw)
    finally:
        ops.conv3x3_halo_policy(1)
    _close(out3, ref.permute(0, 2, 3, 1), f"tap conv {mode} {B}x{H}x{W} {Cin}->{Cout}")


@pytest.mark.parametrize("B,H,W,Cin,Cout,crop", [(2, 16, 16, 64, 64, 0), (3, 32, 32, 640, 640, 0), (2, 32, 66, 320, 320, 1), (1, 8, 22, 1280, 1280, 1),
                                                 (1, 128, 272, 512, 512, 0), (2, 5, 7, 72, 40, 0), (1, 16, 34, 64, 8, 1), (40, 4, 4, 1280, 1280, 0)])
def test_conv_upsample2x_subpixel(B, H, W, Cin, Cout, crop):
    """i360_conv_upsample2x_bf16: nearest x2 upsample + conv3x3 as four 2x2-tap convolutions of the low-resolution tensor with
    pre-summed weights, against fp32 torch interpolate -> conv2d (incl. the circular-halo crop of the panorama branch, partial
    channel blocks, H / W tails).  The pre-summed weights are rounded once to bf16 (<= 2^-9 relative per merged tap)."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + H * 3 + W * 5 + Cin + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g).bfloat16()
    out = ops.conv_upsample2x(x, ops.pack_upsample_conv(w), bias, crop=crop)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, w.float(), bias.float(), padding=1)
    if crop:
        ref = ref[..., 2 * crop:-2 * crop]
    assert out.shape == (B, 2 * H, 2 * (W - 2 * crop), Cout)
    _close(out, ref.permute(0, 2, 3, 1), f"subpixel upsample conv {B}x{H}x{W} {Cin}->{Cout} crop {crop}")
    # and against the materialised path it replaces (same kernels, exact bf16 weights): within the weight-rounding noise
    two = ops.conv3x3(ops.upsample2x(x, pad_in=0), ops.pack_conv3x3(w), bias=bias, crop=2 * crop)
    _close(out, two, "subpixel vs upsample + conv3x3", rtol=1.0 / 64, atol_scale=6e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout,pad_lo,crop", [(2, 16, 16, 64, 64, 1, 0), (3, 32, 32, 320, 320, 1, 0), (2, 32, 72, 320, 320, 1, 1),
                                                        (1, 8, 24, 1280, 1280, 1, 1), (2, 64, 128, 128, 128, 0, 0), (2, 6, 10, 72, 40, 1, 0),
                                                        (1, 256, 512, 128, 128, 0, 0), (40, 8, 8, 1280, 1280, 1, 0)])
def test_conv3x3_stride2_implicit(B, H, W, Cin, Cout, pad_lo, crop):
    """i360_conv3x3_s2_bf16: stride-2 conv through TMA boxes with traversal stride 2 (no im2col), against fp32 torch conv2d:
    symmetric pad 1 (Downsample3D), the VAE's pad (0, 1), the panorama branch's circular halo + crop, tails, partial K blocks;
    and bit-for-bit sane against the im2col + GEMM path it replaces."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + H * 3 + W * 5 + Cin + Cout + pad_lo)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g).bfloat16()
    wp = ops.pack_conv3x3(w)
    out = ops.conv3x3_s2(x, wp, bias, pad_lo=pad_lo, crop=crop)
    xin = x.float().permute(0, 3, 1, 2)
    if pad_lo == 0:
        xin = F.pad(xin, (0, 1, 0, 1))
    ref = F.conv2d(xin, w.float(), bias.float(), stride=2, padding=pad_lo)
    if crop:
        ref = ref[..., crop:-crop]
    assert out.shape == (B, H // 2, W // 2 - 2 * crop, Cout)
    _close(out, ref.permute(0, 2, 3, 1), f"stride-2 conv {B}x{H}x{W} {Cin}->{Cout} pad_lo {pad_lo} crop {crop}")
    if crop == 0:
        two = ops.gemm(ops.im2col_s2(x, False, pad_lo=pad_lo), wp, bias=bias).view(B, H // 2, W // 2, Cout)
        _close(out, two, "implicit vs im2col + GEMM", rtol=1.0 / 128, atol_scale=2e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout,mode", [(2, 64, 136, 128, 128, "bias"), (1, 128, 272, 256, 256, "resid"), (3, 32, 40, 512, 512, "resid"),
                                                 (2, 24, 20, 64, 256, "shortcut"), (2, 16, 24, 8, 512, "bias"), (2, 32, 68, 256, 256, "upsample"),
                                                 (5, 16, 8, 128, 128, "bias")])
def test_conv_with_groupnorm_statistics(B, H, W, Cin, Cout, mode):
    """i360_conv3x3_gnstats_bf16 / i360_conv_upsample2x_gnstats_bf16: the conv result is the plain kernel's and the statistics the
    epilogue accumulated equal those of the separate pass over the stored tensor (to fp32 partial-sum noise); GroupNorm applied
    with either agrees.  Group sizes 4 / 8 / 16, halo and tap-by-tap kernels, residual ring, fused shortcut, tails, 5 images."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + H * 3 + W * 5 + Cin + Cout)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x = rn(B, H, W, Cin).bfloat16()
    w = (rn(Cout, Cin, 3, 3) / (9 * Cin) ** 0.5).bfloat16()
    bias = (rn(Cout) + 0.5).bfloat16()
    kw = {}
    if mode == "upsample":
        weff = ops.pack_upsample_conv(w)
        out, st = ops.conv_upsample2x(x, weff, bias, gn_groups=32)
        plain = ops.conv_upsample2x(x, weff, bias)
    else:
        wp = ops.pack_conv3x3(w)
        if mode == "resid":
            kw["resid"] = rn(B, H, W, Cout).bfloat16()
        if mode == "shortcut":
            x2 = rn(B, H, W, 72).bfloat16()
            wp = ops.pack_conv3x3(w, (rn(Cout, 72, 1, 1) / 72 ** 0.5).bfloat16())
            kw["x2"] = x2
        out, st = ops.conv3x3(x, wp, bias=bias, gn_groups=32, **kw)
        plain = ops.conv3x3(x, wp, bias=bias, **kw)
    assert torch.equal(out, plain)
    gs = Cout // 32
    v = out.double().view(B, -1, 32, gs)
    ref = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], -1)            # statistics of the STORED (bf16) tensor
    n = v.shape[1] * gs
    mean_err = ((st[..., 0] - ref[..., 0]).abs() / n).max().item()
    var = ref[..., 1] / n - (ref[..., 0] / n) ** 2
    var_fused = st[..., 1] / n - (st[..., 0] / n) ** 2
    assert mean_err < 2e-3 and ((var_fused - var).abs() / var).max().item() < 5e-3, (mean_err, ((var_fused - var).abs() / var).max().item())
    gam, bet = (1 + 0.1 * rn(Cout)).bfloat16(), (0.1 * rn(Cout)).bfloat16()
    a = ops.groupnorm(out, gam, bet, 32, 1e-6, True, stats=st)
    b_ = ops.groupnorm(out, gam, bet, 32, 1e-6, True)
    assert (a.float() - b_.float()).abs().max().item() <= 0.04 * b_.float().abs().max().item()


@pytest.mark.parametrize("B,H,W,Cin,Cout,mode", [(2, 64, 64, 320, 320, "temb"), (2, 64, 132, 320, 320, "temb_crop"), (4, 32, 32, 320, 640, "shortcut"),
                                                 (3, 16, 36, 1280, 640, "shortcut_crop"), (5, 8, 8, 1280, 1280, "resid"), (2, 8, 20, 640, 1280, "resid_crop"),
                                                 (7, 4, 8, 1280, 1280, "temb"), (2, 32, 68, 640, 640, "resid")])
def test_conv_with_channel_statistics(B, H, W, Cin, Cout, mode):
    """i360_conv3x3_chanstats_bf16 + i360_groupnorm_apply_chanstats (the UNet's group sizes 10 / 20 / 40): the conv result is
    the plain kernel's, the per-channel sums equal those of the stored tensor, and GroupNorm applied from them agrees with the
    statistics pass.  temb rowvec, fused shortcut, residual, pano crop, several images per tile (TB = 2, 4), halo kernels."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + H * 3 + W * 5 + Cin + Cout)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    crop = 2 if mode.endswith("_crop") else 0
    x = rn(B, H, W, Cin).bfloat16()
    w = (rn(Cout, Cin, 3, 3) / (9 * Cin) ** 0.5).bfloat16()
    bias = (rn(Cout) + 0.5).bfloat16()
    kw = {"crop": crop}
    wp = ops.pack_conv3x3(w)
    if mode.startswith("temb"):
        kw["rowvec"] = rn(1, Cout).float().contiguous()
        kw["rowvec_div"] = B
    if mode.startswith("resid"):
        kw["resid"] = rn(B, H, W - 2 * crop, Cout).bfloat16()
        kw["out_scale"] = 0.5
    if mode.startswith("shortcut"):
        x2 = rn(B, H, W - 2 * crop, 200).bfloat16()
        wp = ops.pack_conv3x3(w, (rn(Cout, 200, 1, 1) / 200 ** 0.5).bfloat16())
        kw["x2"] = x2
    out, st = ops.conv3x3(x, wp, bias=bias, chan_stats=True, **kw)
    plain = ops.conv3x3(x, wp, bias=bias, **kw)
    assert torch.equal(out, plain)
    v = out.double().view(B, -1, Cout)
    ref = torch.stack([v.sum(dim=1), (v * v).sum(dim=1)], -1)                      # statistics of the STORED (bf16) tensor
    n = v.shape[1]
    assert ((st[..., 0] - ref[..., 0]).abs() / n).max().item() < 1e-4
    assert ((st[..., 1] - ref[..., 1]).abs() / ref[..., 1]).max().item() < 1e-4
    gam, bet = (1 + 0.1 * rn(Cout)).bfloat16(), (0.1 * rn(Cout)).bfloat16()
    for silu in (True, False):
        a = ops.groupnorm(out, gam, bet, 32, 1e-5, silu, chan_stats=st)
        b_ = ops.groupnorm(out, gam, bet, 32, 1e-5, silu)
        assert (a.float() - b_.float()).abs().max().item() <= 0.02 * b_.float().abs().max().item()


@pytest.mark.parametrize("M,N,K,resid", [(1000, 320, 320, True), (777, 320, 1280, False), (4100, 640, 640, True), (40000, 640, 640, True),
                                          (38000, 640, 2560, False), (513, 1280, 1280, True), (130, 1280, 5120, False), (64, 320, 320, True)])
def test_gemm_rowstats(M, N, K, resid):
    """i360_gemm_rowstats_bf16: the GEMM result is the plain kernel's, and the slots add up to the row sums / sums of
    squares of what was stored (ring / direct residual epilogues, 160- and 256-wide tiles, the 256 x 128 variant)."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    r = (torch.randn(M, N, device="cuda", generator=g) + 0.5).bfloat16() if resid else None
    out, st = ops.gemm(a, w, bias=b, resid=r, rowstats=True)
    plain = ops.gemm(a, w, bias=b, resid=r)
    assert torch.equal(out, plain)
    ref = a.float() @ w.float().t() + b.float() + (r.float() if resid else 0)
    tot = st.buf.sum(0)
    assert st.buf.shape == (st.slots, M, 2)
    e1 = (tot[:, 0] - ref.sum(1)).abs().max().item()
    e2 = ((tot[:, 1] - (ref * ref).sum(1)).abs() / (ref * ref).sum(1)).max().item()
    assert e1 < 2e-3 * N ** 0.5 and e2 < 1e-4, (e1, e2)
    out2, st2 = ops.gemm(a, w, bias=b, resid=r, rowstats=True)
    assert torch.equal(st.buf, st2.buf), "row statistics must be deterministic (no atomics)"


@pytest.mark.parametrize("M,N,K,act,pe", [(1000, 960, 320, 0, False), (777, 320, 320, 0, False), (513, 2560, 320, 1, False),
                                          (300, 3840, 1280, 0, True), (4100, 1920, 640, 0, True), (129, 1280, 1280, 0, False),
                                          (2000, 5120, 640, 1, False), (64, 960, 320, 0, True), (40000, 640, 640, 0, False),
                                          (3000, 960, 320, 0, 256), (1500, 1920, 640, 0, 128)])
def test_gemm_with_folded_layernorm(M, N, K, act, pe):
    """i360_gemm_ln_bf16: LayerNorm folded into the consuming projection -- row statistics written by the epilogue of the
    GEMM that produced the token matrix, applied in the consumer's epilogue -- against fp32 torch LayerNorm -> Linear
    (-> GEGLU), incl. the temporal-PE row vector."""
    from imagine360_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    # the token matrix x [M, K] comes out of a producer GEMM (bias + residual, like to_out / ff.net[2]) with a row offset
    a0 = torch.randn(M, 320, device="cuda", generator=g).bfloat16()
    w0 = (torch.randn(K, 320, device="cuda", generator=g) * (1.7 / 320 ** 0.5)).bfloat16()
    r0 = (0.6 * torch.randn(M, 1, device="cuda", generator=g)).expand(M, K).contiguous().bfloat16()
    x, st = ops.gemm(a0, w0, resid=r0, rowstats=True)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16() if act else None
    gamma = (1 + 0.2 * torch.randn(K, device="cuda", generator=g)).bfloat16()
    beta = (0.2 * torch.randn(K, device="cuda", generator=g)).bfloat16()
    assert ops.gemm_ln_supported(N, K, act)
    wf, u, c = ops.fold_layernorm(w, b, gamma, beta, geglu=bool(act))
    Fr, D = 4, (5 if pe is True else int(pe))      # D a multiple of 128: the PE vector rides in the tile's bias table
    table = torch.randn(Fr, K, device="cuda", generator=g).bfloat16().float() if pe else None
    rv = (table @ w.float().t()).contiguous() if pe else None
    out = ops.gemm_ln(x, st, wf, u, c, 1e-5, rowvec=rv, rowvec_div=D, rowvec_mod=Fr if pe else 0, act=act)
    y = F.layer_norm(x.float(), (K,), gamma.float(), beta.float(), 1e-5)
    if pe:
        y = y + table[(torch.arange(M, device="cuda") // D) % Fr]
    ref = y @ w.float().t()
    if act:
        ref = ref + b.float()
        val, gate = ref.chunk(2, dim=-1)
        ref = val * F.gelu(gate)
    _close(out, ref, f"gemm_ln {M}x{N}x{K} act{act}")
    # and within a bf16 round-off of the two-kernel path it replaces (which rounds the normalised activations to bf16)
    nrm = ops.layernorm(x, gamma, beta, 1e-5, post_add=table, post_div=D, post_mod=Fr if pe else 1)
    if act:
        wp, bp = ops.pack_geglu(w, b)
        two = ops.gemm(nrm, wp, bias=bp, act=ops.ACT_GEGLU)
    else:
        two = ops.gemm(nrm, w)
    _close(out, two, "folded vs LayerNorm + GEMM", rtol=1.0 / 64, atol_scale=1e-2)
