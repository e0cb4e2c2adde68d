"""Small instances of every kernel path added in round 2, for compute-sanitizer (memcheck / racecheck):
halo-tile conv (all four epilogues), GEMM + row statistics (ring / direct / 256x128), LayerNorm-folded GEMM (plain, GEGLU, PE in
the table, PE per row; two and four epilogue groups via I360_EPI_GROUPS), relative-position bias, per-item attention bias.
Each result is also checked against torch so that a sanitizer-clean run is a correct run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from imagine360_b200 import ops

BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)


def close(a, b, what, tol=2e-2):
    e = ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
    print(f"{what}: rel err {e:.4f}", flush=True)
    assert e < tol, what


# halo conv
ops.conv3x3_halo_policy(1, 1.5, 1, 0)
for (B, H, W, Ci, Co, mode) in [(1, 32, 32, 128, 160, "bias"), (1, 32, 40, 64, 64, "resid"), (2, 16, 24, 72, 128, "rowvec"), (1, 32, 36, 64, 192, "crop"),
                                (1, 16, 24, 64, 64, "shortcut")]:
    x, w, b = rn(B, H, W, Ci).to(BF), (rn(Co, Ci, 3, 3) / (9 * Ci) ** 0.5).to(BF), rn(Co).to(BF)
    crop = 2 if mode == "crop" else 0
    kw, ref = {}, F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=1)
    if crop:
        ref = ref[..., crop:-crop]
    wp = ops.pack_conv3x3(w)
    if mode in ("resid", "crop"):
        kw["resid"] = rn(B, H, W - 2 * crop, Co).to(BF); ref = ref + kw["resid"].float().permute(0, 3, 1, 2)
    if mode == "rowvec":
        kw["rowvec"], kw["rowvec_div"] = rn(B // 2, Co), 2; ref = ref + kw["rowvec"].repeat_interleave(2, 0)[:, :, None, None]
    if mode == "shortcut":
        x2 = rn(B, H, W, 72).to(BF); ws = (rn(Co, 72, 1, 1) / 72 ** 0.5).to(BF); wp = ops.pack_conv3x3(w, ws); kw["x2"] = x2
        ref = ref + F.conv2d(x2.float().permute(0, 3, 1, 2), ws.float())
    assert ops.conv3x3_uses_halo(B, H, W, Ci, "resid" in kw, "rowvec" in kw, "x2" in kw)
    close(ops.conv3x3(x, wp, bias=b, crop=crop, **kw), ref.permute(0, 2, 3, 1), f"halo conv {mode}")
ops.conv3x3_halo_policy(1, 1.04, 0, 64)

# GEMM + row statistics -> folded consumers
for (M, K0, K, N, act, pe) in [(300, 320, 320, 960, 0, 0), (300, 320, 320, 2560, 1, 0), (520, 320, 640, 1920, 0, 128), (300, 320, 1280, 1280, 0, 5),
                               (38000, 640, 640, 640, 0, 0)]:
    a0, w0, r0 = rn(M, K0).to(BF), (rn(K, K0) / K0 ** 0.5).to(BF), rn(M, K).to(BF)
    x, st = ops.gemm(a0, w0, resid=r0, rowstats=True)
    ref_x = a0.float() @ w0.float().t() + r0.float()
    tot = st.buf.sum(0)
    assert (tot[:, 0] - ref_x.sum(1)).abs().max().item() < 0.05 * K ** 0.5, "row sums"
    w = (rn(N, K) / K ** 0.5).to(BF); b = rn(N).to(BF) if act else None
    gam, bet = (1 + 0.2 * rn(K)).to(BF), (0.2 * rn(K)).to(BF)
    wf, u, c = ops.fold_layernorm(w, b, gam, bet, geglu=bool(act))
    Fr = 4
    table = rn(Fr, K).to(BF).float() if pe else None
    rv = (table @ w.float().t()).contiguous() if pe else None
    out = ops.gemm_ln(x, st, wf, u, c, 1e-5, rowvec=rv, rowvec_div=pe or 1, rowvec_mod=Fr if pe else 0, act=act)
    y = F.layer_norm(x.float(), (K,), gam.float(), bet.float(), 1e-5)
    if pe:
        y = y + table[(torch.arange(M, device="cuda") // pe) % Fr]
    ref = y @ w.float().t()
    if act:
        ref = ref + b.float(); v_, g_ = ref.chunk(2, -1); ref = v_ * F.gelu(g_)
    close(out, ref, f"rowstats + folded LN M={M} K={K} N={N} act={act} pe={pe}")

# SAM pieces
for (S, items, heads) in [(14, 3, 2), (3, 2, 2), (16, 1, 3)]:
    hd, N = 64, S * S
    qkv = rn(items * N, 3 * heads * hd).to(BF)
    rh, rw = (0.3 * rn(2 * S - 1, hd)).to(BF), (0.3 * rn(2 * S - 1, hd)).to(BF)
    bias = ops.relpos_bias(qkv, 0, items, heads, hd, S, rh, rw)
    c = heads * hd
    o = torch.empty(items * N, c, device="cuda", dtype=BF)
    ops.attention_item_bias(ops.seq_view(qkv, items, N, 0), ops.seq_view(qkv, items, N, c), ops.seq_view(qkv, items, N, 2 * c),
                            ops.seq_view(o, items, N, 0), heads, hd, items, bias)
    q, k, v = (qkv[:, i * c:(i + 1) * c].float().view(items, N, heads, hd).transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=bias[..., :N].float().view(items, heads, N, N)).transpose(1, 2).reshape(items * N, c)
    close(o, ref, f"relpos bias + item-bias attention S={S}")
torch.cuda.synchronize()
print("all new-kernel instances OK")
