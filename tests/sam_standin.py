"""A stand-in for ``segment_anything.modeling.Sam`` (the package is a dependency of the reference, requirements.txt:11, and
is not installed in this image): parameter containers with the package's module / parameter names and attributes
(``image_encoder.{patch_embed.proj, pos_embed, blocks[i].{norm1, attn.{qkv, proj, rel_pos_h, rel_pos_w, num_heads}, norm2,
mlp.{lin1, lin2}, window_size}, neck[0..3], img_size}``, ``pixel_mean``, ``pixel_std``).  ``forward`` is the oracle
restatement, so the stand-in can also serve CPU plumbing tests.  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn as nn

from oracle import encoders as OE


class LayerNorm2d(nn.Module):
    def __init__(self, c, eps=1e-6):
        super().__init__()
        self.weight, self.bias, self.eps = nn.Parameter(torch.ones(c)), nn.Parameter(torch.zeros(c)), eps


class _Attn(nn.Module):
    def __init__(self, dim, heads, S):
        super().__init__()
        self.num_heads = heads
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)
        self.rel_pos_h = nn.Parameter(torch.zeros(2 * S - 1, dim // heads))
        self.rel_pos_w = nn.Parameter(torch.zeros(2 * S - 1, dim // heads))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.lin1, self.lin2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, window_size, grid):
        super().__init__()
        self.norm1, self.norm2 = nn.LayerNorm(dim, eps=1e-6), nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, heads, window_size if window_size > 0 else grid)
        self.mlp = _Mlp(dim, mlp_ratio * dim)
        self.window_size = window_size


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, patch, stride=patch)


class ImageEncoderViT(nn.Module):
    def __init__(self, embed=768, depth=12, heads=12, patch=16, img=1024, window=14, global_idx=(2, 5, 8, 11), out_chans=256,
                 mlp_ratio=4):
        super().__init__()
        self.img_size, self.heads, self.window, self.global_idx = img, heads, window, tuple(global_idx)
        grid = img // patch
        self.patch_embed = _PatchEmbed(embed, patch)
        self.pos_embed = nn.Parameter(torch.zeros(1, grid, grid, embed))
        self.blocks = nn.ModuleList(_Block(embed, heads, mlp_ratio, 0 if i in global_idx else window, grid) for i in range(depth))
        self.neck = nn.Sequential(nn.Conv2d(embed, out_chans, 1, bias=False), LayerNorm2d(out_chans),
                                  nn.Conv2d(out_chans, out_chans, 3, padding=1, bias=False), LayerNorm2d(out_chans))

    def forward(self, x):
        return OE.sam_image_encoder_forward(self.state_dict(), x, self.heads, self.window, self.global_idx, eps=1e-6)


class Sam(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.image_encoder = ImageEncoderViT(**kw)
        self.register_buffer("pixel_mean", torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1), False)

    def preprocess(self, x):
        return OE.sam_preprocess(x, self.pixel_mean, self.pixel_std, self.image_encoder.img_size)
