"""CPU: oracle/encoders.py (CLIP text encoder, SAM ViT image encoder -- SURVEY.md 8(f) row 2) against the fixtures written
by tools/make_golden_encoders.py from the transformers package the reference itself imports (CLIPTextModel) and from
transformers' port of segment_anything's ImageEncoderViT."""
import pytest
import torch

from golden_util import load, synth_state, synth_tensor
from oracle import encoders as OE


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_clip_text_oracle_matches_transformers_golden(act):
    g = load("encoders.pt")[f"clip_{act}"]
    sd = synth_state(g["shapes"], g["seed"])
    y = OE.clip_text_forward(sd, g["ids"], heads=g["cfg"]["heads"], eps=g["cfg"]["eps"], act=act)
    assert y.shape == g["out"].shape
    assert (y - g["out"]).abs().max().item() < 2e-5


def test_sam_encoder_oracle_matches_hf_port_golden():
    g = load("encoders.pt")["sam"]
    c = g["cfg"]
    sd = synth_state(g["shapes"], g["seed"])
    x = synth_tensor((2, 3, c["img"], c["img"]), g["x_seed"])
    y = OE.sam_image_encoder_forward(sd, x, heads=c["heads"], window_size=c["window"], global_attn_indexes=c["global_idx"], eps=c["eps"])
    assert y.shape == g["out"].shape
    assert (y - g["out"]).abs().max().item() < 5e-5
    # the fixture is not vacuous: the relative position terms and the window padding matter
    sd0 = dict(sd)
    for k in sd0:
        if "rel_pos" in k:
            sd0[k] = torch.zeros_like(sd0[k])
    y0 = OE.sam_image_encoder_forward(sd0, x, heads=c["heads"], window_size=c["window"], global_attn_indexes=c["global_idx"], eps=c["eps"])
    assert (y0 - g["out"]).abs().max().item() > 1e-2
    assert (c["img"] // c["patch"]) % c["window"] != 0


def test_clip_text_oracle_matches_live_transformers():
    tr = pytest.importorskip("transformers")
    g = load("encoders.pt")["clip_gelu"]
    c = g["cfg"]
    cfg = tr.CLIPTextConfig(hidden_size=c["hidden"], intermediate_size=c["inter"], num_hidden_layers=c["layers"],
                            num_attention_heads=c["heads"], vocab_size=c["vocab"], max_position_embeddings=c["max_pos"],
                            hidden_act="gelu", eos_token_id=c["vocab"] - 1, bos_token_id=c["vocab"] - 2, pad_token_id=0)
    m = tr.CLIPTextModel(cfg).eval()
    sd = synth_state(g["shapes"], 21)
    m.load_state_dict(sd, strict=False)
    with torch.no_grad():
        ref = m(g["ids"])[0]
    y = OE.clip_text_forward(sd, g["ids"], heads=c["heads"], eps=c["eps"], act="gelu")
    assert (y - ref).abs().max().item() < 2e-5


def test_sam_key_maps_are_inverse():
    s = OE.sam_shapes()
    assert OE.hf_sam_keys(OE.sa_to_hf_sam_keys(s)) == s
    assert "neck.conv1.weight" in OE.sa_to_hf_sam_keys(s) and "layers.0.layer_norm1.weight" in OE.sa_to_hf_sam_keys(s)


def test_sam_preprocess_pads_bottom_right():
    x = torch.full((1, 3, 4, 8), 10.0)
    y = OE.sam_preprocess(x, torch.tensor([1.0, 2.0, 3.0]), torch.tensor([1.0, 2.0, 4.0]), 8)
    assert y.shape == (1, 3, 8, 8) and y[0, 0, 0, 0] == 9 and y[0, 1, 3, 7] == 4 and (y[:, :, 4:] == 0).all()
