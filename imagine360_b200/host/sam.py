"""The part of ``segment_anything`` the reference touches (requirements.txt:11; inference_dual_p2e.py:369-370
``sam_model_registry["vit_b"](checkpoint=...)``; pipeline_animation_inference_dual.py:171-174 ``SamPredictor``), mirrored so
that a box without the package still runs the script: parameter containers with the package's module / parameter names
(a SAM checkpoint loads by name; the prompt encoder and mask decoder, which the pipeline never runs, are skipped), and the
ViT forward evaluated by :class:`imagine360_b200.host.encoders.SamImageEncoderNative` on the sm_100a kernels.  There is
no CPU path: calling the encoder with a CPU tensor raises.

Shapes follow the published ViT-B / L / H configurations of segment_anything/build_sam.py (release 1.0).
"""
from __future__ import annotations

import sys

import torch
import torch.nn as nn
import torch.nn.functional as F


class LayerNorm2d(nn.Module):
    def __init__(self, num_channels: int, eps: float = 1e-6):
        super().__init__()
        self.weight, self.bias, self.eps = nn.Parameter(torch.ones(num_channels)), nn.Parameter(torch.zeros(num_channels)), eps


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int, input_size: int):
        super().__init__()
        self.num_heads = num_heads
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)
        self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size - 1, dim // num_heads))
        self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size - 1, dim // num_heads))


class MLPBlock(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.lin1, self.lin2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float, window_size: int, grid: int):
        super().__init__()
        self.norm1, self.norm2 = nn.LayerNorm(dim, eps=1e-6), nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads, window_size if window_size > 0 else grid)
        self.mlp = MLPBlock(dim, int(dim * mlp_ratio))
        self.window_size = window_size


class PatchEmbed(nn.Module):
    def __init__(self, dim: int, patch: int):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, patch, stride=patch)


class ImageEncoderViT(nn.Module):
    def __init__(self, img_size=1024, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, out_chans=256,
                 window_size=14, global_attn_indexes=(2, 5, 8, 11)):
        super().__init__()
        self.img_size = img_size
        grid = img_size // patch_size
        self.patch_embed = PatchEmbed(embed_dim, patch_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, grid, grid, embed_dim))
        self.blocks = nn.ModuleList(Block(embed_dim, num_heads, mlp_ratio, 0 if i in global_attn_indexes else window_size, grid)
                                    for i in range(depth))
        self.neck = nn.Sequential(nn.Conv2d(embed_dim, out_chans, 1, bias=False), LayerNorm2d(out_chans),
                                  nn.Conv2d(out_chans, out_chans, 3, padding=1, bias=False), LayerNorm2d(out_chans))

    def forward(self, x):
        if not x.is_cuda:
            msg = "imagine360_b200: the SAM image encoder runs on the CUDA kernels only (no CPU path); move the model and input to a GPU"
            print(msg, file=sys.stderr, flush=True)
            raise RuntimeError(msg)
        from .encoders import SamImageEncoderNative
        nat = self.__dict__.get("_i360_native")
        if nat is None:
            nat = self.__dict__["_i360_native"] = SamImageEncoderNative(self)
        return nat(x)


class Sam(nn.Module):
    """``image_encoder`` + the two normalisation buffers + ``preprocess``: all the pipeline's predictor needs."""
    mask_threshold = 0.0
    image_format = "RGB"

    def __init__(self, image_encoder: ImageEncoderViT, pixel_mean=(123.675, 116.28, 103.53), pixel_std=(58.395, 57.12, 57.375)):
        super().__init__()
        self.image_encoder = image_encoder
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(pixel_std).view(-1, 1, 1), False)

    @property
    def device(self):
        return self.pixel_mean.device

    def preprocess(self, x):
        x = (x - self.pixel_mean) / self.pixel_std
        s = self.image_encoder.img_size
        return F.pad(x, (0, s - x.shape[-1], 0, s - x.shape[-2]))


def _build(embed_dim, depth, num_heads, global_attn_indexes, checkpoint=None):
    sam = Sam(ImageEncoderViT(embed_dim=embed_dim, depth=depth, num_heads=num_heads, global_attn_indexes=tuple(global_attn_indexes)))
    sam.eval()
    if checkpoint is not None:
        with open(checkpoint, "rb") as f:
            sd = torch.load(f, map_location="cpu")
        own = {k: v for k, v in sd.items() if k.startswith("image_encoder.")}      # prompt encoder / mask decoder: never run here
        missing, unexpected = sam.load_state_dict(own, strict=False)
        if missing or unexpected:
            raise RuntimeError(f"imagine360_b200: SAM checkpoint does not match the ViT layout (missing {missing[:3]}, unexpected {unexpected[:3]})")
    return sam


def build_sam_vit_b(checkpoint=None):
    return _build(768, 12, 12, (2, 5, 8, 11), checkpoint)


def build_sam_vit_l(checkpoint=None):
    return _build(1024, 24, 16, (5, 11, 17, 23), checkpoint)


def build_sam_vit_h(checkpoint=None):
    return _build(1280, 32, 16, (7, 15, 23, 31), checkpoint)


sam_model_registry = {"default": build_sam_vit_h, "vit_h": build_sam_vit_h, "vit_l": build_sam_vit_l, "vit_b": build_sam_vit_b}
