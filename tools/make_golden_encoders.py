"""Fixture generator for the conditioning encoders, run in the build container:

    python tools/make_golden_encoders.py        # -> tests/golden/encoders.pt

The reference builds both encoders from third-party packages (inference_dual_p2e.py:369-370,:387): CLIPTextModel from
``transformers`` and SAM ViT-B from ``segment_anything``.  transformers 5.5.0 is installed in this image, so the CLIP
outputs below ARE the dependency's; segment_anything is not, and the SAM outputs come from transformers'
``SamVisionEncoder`` port of the same published algorithm (parameter names mapped by oracle.encoders.sa_to_hf_sam_keys).
Fixtures hold {key: shape}, a seed, the inputs' seeds and the module outputs -- never weights.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import GOLDEN_DIR, synth_state, synth_tensor  # noqa: E402
from oracle import encoders as OE  # noqa: E402

CLIP = dict(hidden=128, inter=256, layers=2, vocab=120, max_pos=77, heads=2)
SAM = dict(embed=128, depth=3, heads=2, patch=8, img=112, window=3, global_idx=(1,), out_chans=32, mlp_ratio=4)


def main():
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel, SamVisionConfig
    from transformers.models.sam.modeling_sam import SamVisionEncoder

    torch.manual_seed(0)
    out = {"transformers": transformers.__version__, "torch": torch.__version__}
    # ---- CLIP text (two activations: SD-2.1's OpenCLIP-H tower uses erf-GELU, OpenAI CLIP quick-GELU) ----
    for act in ("gelu", "quick_gelu"):
        shapes = OE.clip_shapes(**{k: v for k, v in CLIP.items() if k != "heads"})
        sd = synth_state(shapes, 11)
        cfg = CLIPTextConfig(hidden_size=CLIP["hidden"], intermediate_size=CLIP["inter"], num_hidden_layers=CLIP["layers"],
                             num_attention_heads=CLIP["heads"], vocab_size=CLIP["vocab"], max_position_embeddings=CLIP["max_pos"],
                             hidden_act=act, eos_token_id=CLIP["vocab"] - 1, bos_token_id=CLIP["vocab"] - 2, pad_token_id=0)
        m = CLIPTextModel(cfg).eval()
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
        ids = torch.randint(1, CLIP["vocab"] - 2, (2, 77), generator=torch.Generator().manual_seed(5))
        ids[:, 0] = CLIP["vocab"] - 2
        ids[0, 9:] = CLIP["vocab"] - 1        # eos then padding with eos, like the SD-2.1 tokenizer
        ids[1, 1:] = CLIP["vocab"] - 1        # the empty (negative) prompt
        with torch.no_grad():
            y = m(ids)[0]
        out[f"clip_{act}"] = dict(shapes=shapes, seed=11, ids=ids, out=y, cfg=dict(CLIP, act=act, eps=cfg.layer_norm_eps))
    # ---- SAM image encoder ----
    shapes = OE.sam_shapes(**SAM)
    sd = synth_state(shapes, 12)
    # rel_pos tables / pos_embed are zero-initialised in the published model: give them real values
    cfg = SamVisionConfig(hidden_size=SAM["embed"], output_channels=SAM["out_chans"], num_hidden_layers=SAM["depth"],
                          num_attention_heads=SAM["heads"], image_size=SAM["img"], patch_size=SAM["patch"],
                          window_size=SAM["window"], global_attn_indexes=list(SAM["global_idx"]),
                          mlp_dim=SAM["mlp_ratio"] * SAM["embed"])
    m = SamVisionEncoder(cfg).eval()
    missing, unexpected = m.load_state_dict(OE.sa_to_hf_sam_keys(sd), strict=True)
    x = synth_tensor((2, 3, SAM["img"], SAM["img"]), 6)
    with torch.no_grad():
        y = m(x)[0]
    out["sam"] = dict(shapes=shapes, seed=12, x_seed=6, out=y, cfg=dict(SAM, eps=cfg.layer_norm_eps))
    path = os.path.join(GOLDEN_DIR, "encoders.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
