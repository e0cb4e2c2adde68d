"""CPU: the HOST logic of imagine360_b200/host/encoders.py (weight packing, window partition / unpartition, token views, the
order of C-ABI calls) with every kernel entry point replaced by a torch emulation of its documented contract.  The kernels
themselves are checked on the GPU (tests/test_encoders_gpu.py); this keeps index arithmetic mistakes out of GPU time."""
import contextlib

import pytest
import torch
import torch.nn.functional as F

from golden_util import synth_state, synth_tensor
from oracle import encoders as OE

BF16 = torch.bfloat16


def _seq(tv):
    kind, t, n_seq, n_tok, col0, share = tv.src
    assert kind == "seq" and share == 1
    return t, n_seq, n_tok, col0


@contextlib.contextmanager
def emulated_ops():
    from imagine360_b200 import ops
    saved = {k: getattr(ops, k) for k in ("gemm", "layernorm", "attention", "attention_item_bias", "relpos_bias", "conv3x3",
                                          "axpby", "on_device")}

    def gemm(a, w, bias=None, resid=None, act=ops.ACT_NONE, **_):
        y = a.float() @ w.float().t()
        if bias is not None:
            y = y + bias.float()
        if act == ops.ACT_GELU:
            y = F.gelu(y)
        if resid is not None:
            y = y + resid.float()
        return y.to(BF16)

    def layernorm(x, g, b, eps=1e-5, **_):
        return F.layer_norm(x.float(), (x.shape[1],), g.float(), b.float(), eps).to(BF16)

    def _attn(q, k, v, o, heads, hd, batch, bias):
        (qt, ns, n, qc), (kt, _, _, kc), (vt, _, _, vc), (ot, _, _, oc) = _seq(q), _seq(k), _seq(v), _seq(o)
        c = heads * hd
        sp = lambda t, c0: t[:, c0:c0 + c].float().view(ns, n, heads, hd).transpose(1, 2)
        a = sp(qt, qc) @ sp(kt, kc).transpose(-1, -2) * hd ** -0.5 + bias
        ot[:, oc:oc + c] = (a.softmax(-1) @ sp(vt, vc)).transpose(1, 2).reshape(ns * n, c).to(BF16)

    def attention(q, k, v, o, heads, head_dim, batch, scale=None, bias=None, accumulate=False):
        _attn(q, k, v, o, heads, head_dim, batch, bias.float())

    def attention_item_bias(q, k, v, o, heads, head_dim, batch, bias, scale=None):
        n = q.d1
        _attn(q, k, v, o, heads, head_dim, batch, bias[..., :n].float().view(batch, heads, n, n))

    def relpos_bias(qkv, col0, items, heads, hd, S, rel_h, rel_w):
        N = S * S
        q = qkv[:, col0:col0 + heads * hd].float().view(items, S, S, heads, hd).permute(0, 3, 1, 2, 4)
        Rh, Rw = OE._rel_pos(S, S, rel_h.float()), OE._rel_pos(S, S, rel_w.float())
        b = torch.einsum("ihyxc,ykc->ihyxk", q, Rh)[..., :, None] + torch.einsum("ihyxc,xkc->ihyxk", q, Rw)[..., None, :]
        out = torch.zeros((items * heads, N, -(-N // 8) * 8), dtype=BF16)
        out[..., :N] = b.reshape(items * heads, N, N).to(BF16)
        return out

    def conv3x3(x, w_packed, bias=None, **_):
        B, H, W, Cin = x.shape
        w = w_packed.float().view(-1, 3, 3, Cin).permute(0, 3, 1, 2)
        return F.conv2d(x.float().permute(0, 3, 1, 2), w, None, padding=1).permute(0, 2, 3, 1).to(BF16).contiguous()

    def axpby(x, y, a, b):
        return (a * x.float() + b * y.float()).to(BF16)

    for k, f in dict(gemm=gemm, layernorm=layernorm, attention=attention, attention_item_bias=attention_item_bias,
                     relpos_bias=relpos_bias, conv3x3=conv3x3, axpby=axpby, on_device=lambda d: contextlib.nullcontext()).items():
        setattr(ops, k, f)
    try:
        yield
    finally:
        for k, f in saved.items():
            setattr(ops, k, f)


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


@pytest.mark.parametrize("cfg,batch", [
    (dict(embed=128, depth=3, heads=2, patch=8, img=112, window=3, global_idx=(1,), out_chans=32), 2),
    (dict(embed=128, depth=2, heads=2, patch=16, img=256, window=14, global_idx=(0,), out_chans=64), 1),
])
def test_sam_host_logic(cfg, batch):
    from sam_standin import Sam
    from imagine360_b200.host.encoders import SamImageEncoderNative
    sam = Sam(**cfg).eval()
    enc = sam.image_encoder
    sd = synth_state({k: tuple(v.shape) for k, v in enc.state_dict().items()}, 41)
    for k in sd:
        if "rel_pos" in k:
            sd[k] = sd[k] * 4.0
    enc.load_state_dict(sd)
    x = synth_tensor((batch, 3, cfg["img"], cfg["img"]), 7)
    with emulated_ops():
        y = SamImageEncoderNative(enc)(x)
    ref = OE.sam_image_encoder_forward({k: v.to(BF16).float() for k, v in sd.items()}, x.to(BF16).float(), heads=enc.heads,
                                       window_size=enc.window, global_attn_indexes=enc.global_idx)
    assert y.shape == ref.shape and _rel(y, ref) < 3e-2, _rel(y, ref)


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_clip_host_logic(act):
    tr = pytest.importorskip("transformers")
    from imagine360_b200.host.encoders import ClipTextNative
    cfg = tr.CLIPTextConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, vocab_size=120,
                            max_position_embeddings=77, hidden_act=act, eos_token_id=119, bos_token_id=118, pad_token_id=0)
    m = tr.CLIPTextModel(cfg).eval()
    sd = synth_state(OE.clip_shapes(128, 256, 2, 120, 77), 31)
    m.load_state_dict(sd, strict=False)
    ids = torch.randint(1, 118, (2, 77), generator=torch.Generator().manual_seed(5))
    with emulated_ops():
        y = ClipTextNative(m)(ids)[0]
    with torch.no_grad():
        ref = m(ids)[0]
    assert y.shape == ref.shape and _rel(y, ref) < 3e-2, _rel(y, ref)


def test_wrappers_leave_unknown_modules_alone():
    from imagine360_b200.host.encoders import SamImageEncoderNative, wrap_text_encoder
    lin = torch.nn.Linear(4, 4)
    assert wrap_text_encoder(lin) is lin and wrap_text_encoder(None) is None
    assert not SamImageEncoderNative.supports(lin)
    with pytest.raises(TypeError):
        SamImageEncoderNative(lin)


def test_segment_anything_mirror_layout_and_dropin():
    """inference_dual_p2e.py:369-370 `from segment_anything import SamPredictor, sam_model_registry` resolves through the
    drop-in; the ViT-B container has the published layout (89.67 M parameters, segment_anything parameter names) and no
    CPU forward."""
    import sys
    import imagine360_b200.dropin as dropin
    had = sys.modules.get("segment_anything")
    try:
        done = dropin.install()
        assert set(done["segment_anything"]) == {"sam_model_registry", "SamPredictor"}
        from segment_anything import SamPredictor, sam_model_registry
        sam = sam_model_registry["vit_b"]()
        enc = sam.image_encoder
        n = sum(p.numel() for p in enc.parameters())
        assert n == 89_670_912, n
        sd = enc.state_dict()
        assert sd["blocks.2.attn.rel_pos_h"].shape == (127, 64) and sd["blocks.0.attn.rel_pos_h"].shape == (27, 64)
        assert sd["pos_embed"].shape == (1, 64, 64, 768) and sd["neck.2.weight"].shape == (256, 256, 3, 3)
        assert [b.window_size for b in enc.blocks] == [14, 14, 0, 14, 14, 0, 14, 14, 0, 14, 14, 0]
        pred = SamPredictor(sam)
        assert pred.native_encoder is not None and pred.transform.target_length == 1024
        with pytest.raises(RuntimeError):
            enc(torch.zeros(1, 3, 1024, 1024))
    finally:
        if had is None:
            sys.modules.pop("segment_anything", None)
        else:
            sys.modules["segment_anything"] = had
