"""Test-side SAM models: the product's ``segment_anything`` mirror (imagine360_b200/host/sam.py: parameter containers with
the package's names, native forward, no CPU path) with the ORACLE as forward, so CPU plumbing tests can run them and GPU
tests have an fp32 torch module of the same parameters.  TEST INFRASTRUCTURE ONLY."""
import torch

from imagine360_b200.host import sam as S
from oracle import encoders as OE


class ImageEncoderViT(S.ImageEncoderViT):
    def __init__(self, embed=768, depth=12, heads=12, patch=16, img=1024, window=14, global_idx=(2, 5, 8, 11), out_chans=256,
                 mlp_ratio=4):
        super().__init__(img_size=img, patch_size=patch, embed_dim=embed, depth=depth, num_heads=heads, mlp_ratio=mlp_ratio,
                         out_chans=out_chans, window_size=window, global_attn_indexes=tuple(global_idx))
        self.heads, self.window, self.global_idx = heads, window, tuple(global_idx)

    def forward(self, x):
        return OE.sam_image_encoder_forward(self.state_dict(), x, self.heads, self.window, self.global_idx, eps=1e-6)


class Sam(S.Sam):
    def __init__(self, **kw):
        super().__init__(ImageEncoderViT(**kw))
