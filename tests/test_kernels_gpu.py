"""GPU parity of the non-GEMM kernels (norms, attention variants, temporal attention, data movement,
CFG+DDIM) against fp32 torch math on the same bf16 inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(out, ref, what, rtol=1.0 / 128, atol_scale=3e-3):
    out, ref = out.float(), ref.float()
    assert out.shape == ref.shape, (what, out.shape, ref.shape)
    atol = atol_scale * ref.abs().max().item() + 1e-6
    err = (out - ref).abs()
    bad = err > (atol + rtol * ref.abs())
    assert not bad.any(), f"{what}: {bad.sum().item()}/{bad.numel()} mismatches, max err {err.max().item():.4g} (ref max {ref.abs().max().item():.4g})"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).bfloat16()


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,C1,C2,pad,groups,silu", [
    (3, 8, 8, 320, 0, 0, 32, True), (2, 16, 20, 640, 320, 2, 32, True), (5, 4, 4, 1280, 1280, 0, 32, True),
    (2, 6, 10, 64, 0, 1, 32, False), (1, 64, 136, 128, 0, 0, 32, True), (2, 8, 16, 960, 0, 2, 32, True),
])
def test_groupnorm(B, H, W, C1, C2, pad, groups, silu):
    from imagine360_b200 import ops
    x1 = _rand(B, H, W, C1, seed=1) + 0.5
    x2 = _rand(B, H, W, C2, seed=2) * 2 if C2 else None
    C = C1 + C2
    gamma, beta = _rand(C, seed=3) * 0.2 + 1, _rand(C, seed=4) * 0.2
    out = ops.groupnorm(x1, gamma, beta, groups, 1e-5, silu, x2=x2, pad=pad)
    xc = torch.cat([x1, x2], -1) if C2 else x1
    xn = xc.float().permute(0, 3, 1, 2)
    if pad:
        xn = torch.cat([xn[..., -pad:], xn, xn[..., :pad]], -1)
    ref = F.group_norm(xn, groups, gamma.float(), beta.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    _close(out, ref.permute(0, 2, 3, 1), "groupnorm")


def test_groupnorm_stats_unpadded_output_padded():
    """conv_norm_out: statistics before the circular pad, output written with the pad (MVGenModel.py:472-475)."""
    from imagine360_b200 import ops
    x = _rand(2, 8, 16, 320, seed=5)
    gamma, beta = _rand(320, seed=6) * 0.2 + 1, _rand(320, seed=7) * 0.2
    out = ops.groupnorm(x, gamma, beta, 32, 1e-5, True, pad=1, stats_pad=0)
    ref = F.silu(F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma.float(), beta.float(), 1e-5))
    ref = torch.cat([ref[..., -1:], ref, ref[..., :1]], -1).permute(0, 2, 3, 1)
    _close(out, ref, "groupnorm pad-after")


@pytest.mark.parametrize("M,C", [(1000, 320), (77, 1024), (513, 1280), (64, 640)])
def test_layernorm(M, C):
    from imagine360_b200 import ops
    x = _rand(M, C, seed=8) * 2 + 0.3
    gamma, beta = _rand(C, seed=9) * 0.2 + 1, _rand(C, seed=10) * 0.2
    _close(ops.layernorm(x, gamma, beta), F.layer_norm(x.float(), (C,), gamma.float(), beta.float()), "layernorm")
    # temporal PE added after (rows ordered (b f d): frame = (row // D) % F)
    Fr, D = 4, 5
    pe = torch.randn(Fr, C, device="cuda")
    rows = torch.arange(M, device="cuda")
    ref = F.layer_norm(x.float(), (C,), gamma.float(), beta.float()).bfloat16().float() + pe[(rows // D) % Fr]
    _close(ops.layernorm(x, gamma, beta, post_add=pe, post_div=D, post_mod=Fr), ref, "layernorm+post")
    # spherical PE added before, view-major table: row -> ((row // 12) % 3) * 4 + row % 4
    tab = _rand(12, C, seed=11)
    idx = ((rows // 12) % 3) * 4 + rows % 4
    ref = F.layer_norm((x + tab[idx]).float(), (C,), gamma.float(), beta.float())
    _close(ops.layernorm(x, gamma, beta, pre_add=tab, pre_index=(12, 3, 4, 4)), ref, "layernorm+pre")


# ------------------------------------------------------------------------------------------------
def _sdpa(q, k, v, heads, bias=None):
    b, n, c = q.shape
    d = c // heads
    qh, kh, vh = (t.float().reshape(t.shape[0], t.shape[1], heads, d).transpose(1, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=None if bias is None else bias.float())
    return o.transpose(1, 2).reshape(b, n, c)


@pytest.mark.parametrize("imgs,N,heads,hd", [(3, 256, 5, 64), (2, 1024, 2, 64), (4, 64, 3, 64), (5, 16, 2, 64),
                                             (2, 200, 2, 64), (1, 2048, 1, 64), (3, 128, 4, 32), (2, 300, 3, 32)])
def test_self_attention_fused_qkv(imgs, N, heads, hd):
    """spatial attn1: q, k, v are column slices of one fused projection output [tokens, 3C]."""
    from imagine360_b200 import ops
    C = heads * hd
    qkv = _rand(imgs * N, 3 * C, seed=12)
    out = torch.zeros(imgs * N, C, device="cuda", dtype=torch.bfloat16)
    ops.attention(ops.seq_view(qkv, imgs, N, 0), ops.seq_view(qkv, imgs, N, C), ops.seq_view(qkv, imgs, N, 2 * C),
                  ops.seq_view(out, imgs, N), heads, hd, imgs)
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(imgs, N, C) for i in range(3))
    _close(out.reshape(imgs, N, C), _sdpa(q, k, v, heads), f"self-attn N={N} hd={hd}")


def test_cross_attention_shared_kv_and_accumulate():
    """attn2: K/V rows of a clip are shared by its F frames; the IP branch is summed into the text branch."""
    from imagine360_b200 import ops
    clips, Fr, N, heads, hd = 2, 3, 200, 5, 64
    C = heads * hd
    q = _rand(clips * Fr * N, C, seed=13)
    kv_t = _rand(clips * 77, 2 * C, seed=14)
    kv_i = _rand(clips * 64, 2 * C, seed=15)
    out = torch.empty(clips * Fr * N, C, device="cuda", dtype=torch.bfloat16)
    qv, ov = ops.seq_view(q, clips * Fr, N), ops.seq_view(out, clips * Fr, N)
    ops.attention(qv, ops.seq_view(kv_t, clips, 77, 0, share_div=Fr), ops.seq_view(kv_t, clips, 77, C, share_div=Fr), ov,
                  heads, hd, clips * Fr)
    ops.attention(qv, ops.seq_view(kv_i, clips, 64, 0, share_div=Fr), ops.seq_view(kv_i, clips, 64, C, share_div=Fr), ov,
                  heads, hd, clips * Fr, accumulate=True)
    q3 = q.reshape(clips * Fr, N, C)

    def rep(t, n):
        return t.reshape(clips, n, 2 * C).repeat_interleave(Fr, 0)

    kt, ki = rep(kv_t, 77), rep(kv_i, 64)
    ref = _sdpa(q3, kt[..., :C], kt[..., C:], heads).bfloat16().float() + _sdpa(q3, ki[..., :C], ki[..., C:], heads).bfloat16().float()
    _close(out.reshape(clips * Fr, N, C), ref, "cross-attn text+ip", atol_scale=8e-3)


@pytest.mark.parametrize("clips,Fr,N,heads,nt,ni", [
    (2, 3, 200, 5, 77, 64),      # rows per element not a multiple of 128 (masked tail tile)
    (3, 16, 16, 2, 77, 64),      # mid-block shape: 256 rows per element
    (4, 16, 128, 5, 77, 64),     # > 148 items: several tiles per CTA, ranges crossing (element, head) boundaries
    (6, 16, 256, 10, 77, 64),    # ~13 tiles per CTA
    (2, 8, 64, 3, 77, 16), (2, 8, 64, 2, 20, 4), (1, 8, 96, 2, 96, 64), (2, 2, 24, 1, 1, 1),
])
def test_cross_attention_text_ip_fused(clips, Fr, N, heads, nt, ni):
    """attn2 as ONE kernel with stationary K/V: softmax(q Kt^T) Vt + softmax(q Ki^T) Vi (attention.py:119-148)."""
    from imagine360_b200 import ops
    hd = 64
    C = heads * hd
    assert ops.cross_attention_text_ip_supported(hd, nt, ni)
    q = _rand(clips * Fr * N, C, seed=23)
    kv_t = _rand(clips * nt, 2 * C, seed=24)
    kv_i = _rand(clips * ni, 2 * C, seed=25)
    out = torch.full((clips * Fr * N, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.cross_attention_text_ip(q, out, kv_t, nt, kv_i, ni, clips, heads, hd)
    q3 = q.reshape(clips * Fr, N, C)

    def rep(t, n):
        return t.reshape(clips, n, 2 * C).repeat_interleave(Fr, 0)

    kt, ki = rep(kv_t, nt), rep(kv_i, ni)
    ref = _sdpa(q3, kt[..., :C], kt[..., C:], heads) + _sdpa(q3, ki[..., :C], ki[..., C:], heads)
    _close(out.reshape(clips * Fr, N, C), ref, f"fused cross-attn nt={nt} ni={ni}", atol_scale=8e-3)
    # and against the generic kernel's two-call path (bf16 rounding of each branch before the sum)
    out2 = torch.empty_like(out)
    qv, ov = ops.seq_view(q, clips * Fr, N), ops.seq_view(out2, clips * Fr, N)
    ops.attention(qv, ops.seq_view(kv_t, clips, nt, 0, share_div=Fr), ops.seq_view(kv_t, clips, nt, C, share_div=Fr), ov,
                  heads, hd, clips * Fr)
    ops.attention(qv, ops.seq_view(kv_i, clips, ni, 0, share_div=Fr), ops.seq_view(kv_i, clips, ni, C, share_div=Fr), ov,
                  heads, hd, clips * Fr, accumulate=True)
    _close(out, out2, "fused vs two-call cross-attn", rtol=1.0 / 64, atol_scale=1.2e-2)


def test_cross_attention_text_ip_limits():
    from imagine360_b200 import ops
    assert not ops.cross_attention_text_ip_supported(32, 77, 64)
    assert not ops.cross_attention_text_ip_supported(64, 128, 64)
    q = _rand(256, 32, seed=1)
    with pytest.raises(RuntimeError):
        ops.cross_attention_text_ip(q, torch.empty_like(q), _rand(77, 64, seed=2), 77, _rand(64, 64, seed=3), 64, 1, 1, 32)


@pytest.mark.parametrize("b,m,Fr,ph,eh,ew,heads", [(2, 3, 2, 4, 8, 16, 2), (1, 20, 2, 8, 16, 32, 10), (1, 4, 1, 16, 32, 64, 10),
                                                  (2, 2, 2, 2, 4, 8, 4), (1, 4, 1, 6, 12, 24, 2)])
def test_warp_attention_with_bias(b, m, Fr, ph, eh, ew, heads):
    """WarpAttn: pers tokens live as [(b m f), hw, C]; the sequence of batch item (b, f) gathers the m views."""
    from imagine360_b200 import ops
    hd = 32
    C = heads * hd
    hw, EN = ph * ph, eh * ew
    pers = _rand(b * m * Fr * hw, C, seed=16)          # rows (b, m, f, t)
    pers_kv = _rand(b * m * Fr * hw, 2 * C, seed=17)
    equi = _rand(b * Fr * EN, C, seed=18)              # rows (b, f, t)
    equi_kv = _rand(b * Fr * EN, 2 * C, seed=19)
    bias_e = _rand(EN, m * hw, seed=20)                # [(eh ew), (m ph pw)]
    bias_p = _rand(m * hw, EN, seed=21)
    out_e = torch.zeros_like(equi)
    out_p = torch.zeros_like(pers)
    # equi queries <- pers keys/values
    ops.attention(ops.seq_view(equi, b * Fr, EN), ops.multiview_view(pers_kv, b, m, Fr, hw, 0),
                  ops.multiview_view(pers_kv, b, m, Fr, hw, C), ops.seq_view(out_e, b * Fr, EN), heads, hd, b * Fr, bias=bias_e)
    # pers queries <- equi keys/values
    ops.attention(ops.multiview_view(pers, b, m, Fr, hw), ops.seq_view(equi_kv, b * Fr, EN, 0), ops.seq_view(equi_kv, b * Fr, EN, C),
                  ops.multiview_view(out_p, b, m, Fr, hw), heads, hd, b * Fr, bias=bias_p)

    def to_bf(t, c):   # (b m f t) c -> (b f) (m t) c
        return t.reshape(b, m, Fr, hw, c).permute(0, 2, 1, 3, 4).reshape(b * Fr, m * hw, c)

    pkv = to_bf(pers_kv, 2 * C)
    e3 = equi.reshape(b * Fr, EN, C)
    ref_e = _sdpa(e3, pkv[..., :C], pkv[..., C:], heads, bias_e)
    _close(out_e.reshape(b * Fr, EN, C), ref_e, "warp equi<-pers")
    ekv = equi_kv.reshape(b * Fr, EN, 2 * C)
    ref_p = _sdpa(to_bf(pers, C), ekv[..., :C], ekv[..., C:], heads, bias_p)
    _close(to_bf(out_p, C), ref_p, "warp pers<-equi")


@pytest.mark.parametrize("B,Fr,D,heads,hd", [(2, 16, 50, 8, 40), (1, 16, 33, 8, 80), (2, 8, 20, 8, 160), (1, 24, 10, 8, 40),
                                            (2, 16, 7, 8, 64), (1, 4, 5, 2, 16), (2, 24, 333, 8, 80), (1, 24, 57, 8, 160),
                                            (1, 32, 40, 8, 40), (2, 17, 21, 5, 64)])
def test_temporal_attention(B, Fr, D, heads, hd):
    from imagine360_b200 import ops
    C = heads * hd
    qkv = _rand(B * Fr * D, 3 * C, seed=22)
    out = torch.zeros(B * Fr * D, C, device="cuda", dtype=torch.bfloat16)
    ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, B, Fr, D, heads, hd)

    def bd(t):  # (b f d) c -> (b d) f c
        return t.reshape(B, Fr, D, C).permute(0, 2, 1, 3).reshape(B * D, Fr, C)

    ref = _sdpa(bd(qkv[:, :C]), bd(qkv[:, C:2 * C]), bd(qkv[:, 2 * C:]), heads)
    _close(bd(out), ref, "temporal attention")


# ------------------------------------------------------------------------------------------------
def test_upsample_and_im2col():
    from imagine360_b200 import ops
    x = _rand(2, 6, 10, 64, seed=23)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    upc = ops.upsample2x(x, pad_in=1)
    xp = torch.cat([x[:, :, -1:], x, x[:, :, :1]], 2)
    refc = F.interpolate(xp.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(upc.float(), refc)
    # stride-2 conv through im2col + GEMM: plain (pers) and circular (pano: pad 2 -> conv -> crop 1)
    w = _rand(128, 64, 3, 3, seed=24, scale=0.05)
    bias = _rand(128, seed=25)
    wp = ops.pack_conv3x3(w)
    y = ops.gemm(ops.im2col_s2(x, circular=False), wp, bias=bias).reshape(2, 3, 5, 128)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    _close(y, ref, "downsample")
    yc = ops.gemm(ops.im2col_s2(x, circular=True), wp, bias=bias).reshape(2, 3, 5, 128)
    xp2 = torch.cat([x[:, :, -2:], x, x[:, :, :2]], 2)
    refc = F.conv2d(xp2.float().permute(0, 3, 1, 2), w.float(), bias.float(), stride=2, padding=1)[..., 1:-1].permute(0, 2, 3, 1)
    _close(yc, refc, "pano downsample")


def test_cfg_ddim_bit_exact_vs_torch_bf16():
    """The fused kernel reproduces the reference's bf16 op-by-op rounding exactly."""
    from imagine360_b200 import ops
    x, u, c = _rand(1, 4, 16, 8, 16, seed=26), _rand(1, 4, 16, 8, 16, seed=27), _rand(1, 4, 16, 8, 16, seed=28)
    sa, sb, sap, sbp = (torch.tensor(v, dtype=torch.float32) for v in (0.6123, 0.7906, 0.6541, 0.7564))
    out = ops.cfg_ddim_step(x, u, c, 7.5, float(sa), float(sb), float(sap), float(sbp))
    v = u + 7.5 * (c - u)
    x0 = sa * x - sb * v
    eps = sa * v + sb * x
    ref = sap * x0 + sbp * eps
    assert ref.dtype == torch.bfloat16
    assert torch.equal(out, ref)


def test_axpby_avgpool_gridsample():
    from imagine360_b200 import ops
    x, n = _rand(40, 64, 1024, seed=29), _rand(40, 64, 1024, seed=30)
    assert torch.equal(ops.axpby(x, n, 1.0, 0.1), x + n * 0.1)
    t = _rand(2, 16, 9, 64, seed=31)
    ref = F.avg_pool1d(t.float().permute(0, 2, 3, 1).reshape(-1, 64, 16).reshape(2 * 9 * 64, 1, 16), 4)
    ref = ref.reshape(2, 9, 64, 4).permute(0, 3, 1, 2)
    _close(ops.avgpool_frames4(t), ref, "avgpool", rtol=1 / 256, atol_scale=1e-3)
    img = torch.randn(3, 4, 8, 16, device="cuda")
    grid = torch.rand(3, 5, 7, 2, device="cuda") * 2.4 - 1.2
    for nearest in (False, True):
        ref = F.grid_sample(img, grid, mode="nearest" if nearest else "bilinear", padding_mode="zeros", align_corners=True)
        out = ops.grid_sample(img, grid, nearest=nearest)
        assert (out - ref).abs().max().item() < 1e-5, nearest
