"""GPU parity of SURVEY.md 8(f) row 2: the CLIP text encoder and the SAM ViT image encoder evaluated through the C ABI
against oracle/encoders.py (pinned to transformers by tests/test_encoders_oracle.py), with the calibrated bound of section
8(c): err(native vs fp32 oracle) <= k * err(bf16 torch vs fp32 oracle) + atol, the bf16 torch run being the same oracle on
bf16 tensors (what the reference's torch path would do in the UNets' dtype)."""
import pytest
import torch

from golden_util import synth_state, synth_tensor
from oracle import encoders as OE

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16 = torch.bfloat16


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


@pytest.mark.parametrize("S,items,heads", [(14, 5, 2), (3, 4, 2), (64, 1, 2), (8, 3, 12)])
def test_relpos_bias_vs_einsum(S, items, heads):
    from imagine360_b200 import ops
    hd, N = 64, S * S
    qkv = synth_tensor((items * N, 3 * heads * hd), 1).to(DEV, BF16)
    rh, rw = synth_tensor((2 * S - 1, hd), 2, 0.3).to(DEV, BF16), synth_tensor((2 * S - 1, hd), 3, 0.3).to(DEV, BF16)
    bias = ops.relpos_bias(qkv, 0, items, heads, hd, S, rh, rw)
    torch.cuda.synchronize()
    q = qkv[:, : heads * hd].float().view(items, S, S, heads, hd).permute(0, 3, 1, 2, 4)        # [i, h, qh, qw, c]
    Rh, Rw = OE._rel_pos(S, S, rh.float().cpu()).to(DEV), OE._rel_pos(S, S, rw.float().cpu()).to(DEV)
    ref = (torch.einsum("ihyxc,ykc->ihyxk", q, Rh)[..., :, None] + torch.einsum("ihyxc,xkc->ihyxk", q, Rw)[..., None, :])
    ref = ref.reshape(items * heads, N, N)
    assert bias.shape[:2] == (items * heads, N) and bias.shape[2] % 8 == 0
    assert (bias[..., N:] == 0).all()
    err = (bias[..., :N].float() - ref).abs().max().item()
    assert err <= 2 ** -8 * ref.abs().max().item() + 1e-6, err            # one bf16 rounding of an fp32 result


@pytest.mark.parametrize("N,items,heads", [(196, 6, 2), (9, 4, 2), (1024, 2, 3), (4096, 1, 2)])
def test_attention_item_bias_vs_sdpa(N, items, heads):
    from imagine360_b200 import ops
    hd, c = 64, heads * 64
    qkv = synth_tensor((items * N, 3 * c), 4).to(DEV, BF16)
    ldb = -(-N // 8) * 8
    bias = torch.zeros((items * heads, N, ldb), dtype=BF16, device=DEV)
    bias[..., :N] = synth_tensor((items * heads, N, N), 5, 1.5).to(DEV, BF16)
    o = torch.empty((items * N, c), dtype=BF16, device=DEV)
    ops.attention_item_bias(ops.seq_view(qkv, items, N, 0), ops.seq_view(qkv, items, N, c), ops.seq_view(qkv, items, N, 2 * c),
                            ops.seq_view(o, items, N, 0), heads, hd, items, bias)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * c:(i + 1) * c].float().view(items, N, heads, hd).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias[..., :N].float().view(items, heads, N, N))
    ref = ref.transpose(1, 2).reshape(items * N, c)
    assert _rel(o, ref) < 1.5e-2        # bf16 P and bf16 output on values of order 1


def _clip(hidden, inter, layers, heads, vocab=120, act="gelu"):
    tr = pytest.importorskip("transformers")
    cfg = tr.CLIPTextConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=heads,
                            vocab_size=vocab, max_position_embeddings=77, hidden_act=act, eos_token_id=vocab - 1,
                            bos_token_id=vocab - 2, pad_token_id=0)
    m = tr.CLIPTextModel(cfg).eval()
    sd = synth_state(OE.clip_shapes(hidden, inter, layers, vocab, 77), 31)
    m.load_state_dict(sd, strict=False)
    return m, sd


@pytest.mark.parametrize("hidden,inter,layers,heads,act", [(128, 256, 2, 2, "gelu"), (128, 256, 2, 2, "quick_gelu"),
                                                            (1024, 4096, 23, 16, "gelu")])
def test_clip_text_native_vs_oracle(hidden, inter, layers, heads, act):
    """The last case is the SD-2.1 text tower (OpenCLIP ViT-H: 1024 wide, 23 layers, 16 heads, erf-GELU)."""
    from imagine360_b200 import _lib
    from imagine360_b200.host.encoders import wrap_text_encoder
    m, sd = _clip(hidden, inter, layers, heads, act=act)
    ids = torch.randint(1, 118, (2, 77), generator=torch.Generator().manual_seed(5))
    ids[:, 0], ids[0, 12:], ids[1, 1:] = 118, 119, 119
    m = m.to(DEV)
    enc = wrap_text_encoder(m)
    n0 = _lib.LAUNCHES
    y = enc(ids.to(DEV))[0]
    torch.cuda.synchronize()
    assert _lib.LAUNCHES - n0 >= 7 * layers + 1, "the CUDA path did not run"
    assert y.shape == (2, 77, hidden) and y.dtype == torch.float32
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    ref = OE.clip_text_forward({k: v.to(BF16).float() for k, v in sdd.items()}, ids.to(DEV), heads, act=act)
    lib = OE.clip_text_forward({k: v.to(BF16) for k, v in sdd.items()}, ids.to(DEV), heads, act=act)
    e_nat, e_lib = _rel(y, ref), _rel(lib, ref)
    print(f"clip {hidden}x{layers} {act}: native {e_nat:.4f}  bf16 torch {e_lib:.4f}")
    assert e_nat <= 2.0 * e_lib + 4e-3, (e_nat, e_lib)
    # and against the transformers module itself (fp32, unrounded weights): the end-to-end deviation a user sees
    with torch.no_grad():
        full = m(ids.to(DEV))[0]
    assert _rel(y, full) < 4e-2


def _sam(**kw):
    from sam_standin import Sam
    sam = Sam(**kw).eval()
    enc = sam.image_encoder
    shapes = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    sd = synth_state(shapes, 41)
    for k in sd:
        if "rel_pos" in k:
            sd[k] = sd[k] * 4.0          # make the relative position term matter (entries ~ 0.5)
    enc.load_state_dict(sd)
    return sam, sd


@pytest.mark.parametrize("cfg,batch", [
    (dict(embed=128, depth=3, heads=2, patch=8, img=112, window=3, global_idx=(1,), out_chans=32), 2),
    (dict(embed=128, depth=2, heads=2, patch=16, img=256, window=14, global_idx=(1,), out_chans=64), 3),
    (dict(), 2),          # ViT-B as sam_model_registry["vit_b"] builds it: 768 x 12, 1024 px, 14 x 14 windows, 4 global blocks
])
def test_sam_encoder_native_vs_oracle(cfg, batch):
    from imagine360_b200 import _lib
    from imagine360_b200.host.encoders import SamImageEncoderNative
    sam, sd = _sam(**cfg)
    enc = sam.image_encoder.to(DEV)
    S = enc.img_size
    x = synth_tensor((batch, 3, S, S), 7).to(DEV)
    nat = SamImageEncoderNative(enc)
    n0 = _lib.LAUNCHES
    y = nat(x)
    torch.cuda.synchronize()
    assert _lib.LAUNCHES > n0
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    kw = dict(heads=enc.heads, window_size=enc.window, global_attn_indexes=enc.global_idx, eps=1e-6)
    ref = OE.sam_image_encoder_forward({k: v.to(BF16).float() for k, v in sdd.items()}, x.to(BF16).float(), **kw)
    lib = OE.sam_image_encoder_forward({k: v.to(BF16) for k, v in sdd.items()}, x.to(BF16), **kw)
    assert y.shape == ref.shape
    e_nat, e_lib = _rel(y, ref), _rel(lib, ref)
    print(f"sam {cfg or 'vit_b'}: native {e_nat:.4f}  bf16 torch {e_lib:.4f}")
    assert e_nat <= 2.0 * e_lib + 4e-3, (e_nat, e_lib)


def test_sam_predictor_takes_the_native_encoder_and_matches_the_pipeline_layout():
    """pipeline...dual.py:675-718 through the pipeline's own predictor: uint8 frames -> apply_image -> batches of 8 ->
    get_image_embedding -> 'f c h w -> f (h w) c'."""
    from imagine360_b200 import _lib
    from imagine360_b200.host.pipeline import AnimationPipeline
    sam, sd = _sam(embed=128, depth=2, heads=2, patch=16, img=256, window=14, global_idx=(1,), out_chans=64)
    sam = sam.to(DEV)
    pipe = AnimationPipeline(None, None, None, None, None, None, None, image_encoder=sam, image_encoder_name="SAM")
    pipe.device = torch.device(DEV)
    assert pipe.SAMpredictor.native_encoder is not None
    anchor = (synth_tensor((8, 3, 64, 128), 9) * 0.4).clamp(-1, 1).to(DEV)
    n0 = _lib.LAUNCHES
    feats = pipe._sam_features(anchor)
    torch.cuda.synchronize()
    assert _lib.LAUNCHES > n0 and feats.shape == (8, 16 * 16, 64)
    # the same frames through the oracle (fp32): resize with the predictor's own transform, preprocess, encode
    import numpy as np
    imgs = np.uint8(((anchor.float() + 1.0) / 2.0 * 255).cpu().numpy().transpose(0, 2, 3, 1))
    fr = torch.stack([torch.as_tensor(pipe.SAMProcessor.apply_image(np.ascontiguousarray(i))) for i in imgs]).permute(0, 3, 1, 2).float()
    xin = OE.sam_preprocess(fr, sam.pixel_mean.cpu().flatten(), sam.pixel_std.cpu().flatten(), 256)
    enc = sam.image_encoder
    ref = OE.sam_image_encoder_forward({k: v.to(BF16).float() for k, v in sd.items()}, xin, heads=enc.heads, window_size=enc.window,
                                       global_attn_indexes=enc.global_idx)
    ref = ref.flatten(2).transpose(1, 2)
    assert _rel(feats.cpu(), ref) < 3e-2
