"""Import the UNMODIFIED reference (/root/reference) in this container.

Only used to pin the oracle and to generate the golden fixtures under tests/golden/ (the reference
does not exist on the GPU box).  Recipe follows SURVEY.md §8(c): the reference is pure Python and
imports cleanly once the third-party modules that are not installed here are stood in for:

* ``xformers.ops.memory_efficient_attention`` -> ``F.scaled_dot_product_attention`` (same math:
  softmax(q k^T / sqrt(d) + bias) v on [B*H, N, d] inputs),
* ``kornia`` (unpinned, unvendored): ``remap`` = pixel->[-1,1] normalisation in the grid's own dtype
  followed by ``F.grid_sample(padding_mode='zeros', align_corners=True)``; ``gaussian_blur2d`` =
  separable normalised Gaussian with replicate border; ``create_meshgrid`` in pixel coordinates,
* ``fairscale.nn.checkpoint.checkpoint_wrapper`` = identity, plus inert stubs for imageio / decord /
  geocalib / omegaconf / segment_anything / accelerate bits the hot path never touches.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import math
import sys
import types

import torch
import torch.nn.functional as F

REF = "/root/reference"
_done = False


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    sys.modules[name] = m
    return m


def _kornia():
    def remap(image, map_x, map_y, mode="bilinear", padding_mode="zeros", align_corners=None,
              normalized_coordinates=False):
        b, c, h, w = image.shape
        gx = 2.0 * map_x / (w - 1) - 1.0 if not normalized_coordinates else map_x
        gy = 2.0 * map_y / (h - 1) - 1.0 if not normalized_coordinates else map_y
        grid = torch.stack([gx, gy], dim=-1).to(image.dtype)
        return F.grid_sample(image, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)

    def gaussian_kernel1d(ksize, sigma, device, dtype):
        x = torch.arange(ksize, device=device, dtype=dtype) - ksize // 2
        if ksize % 2 == 0:
            x = x + 0.5
        g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
        return g / g.sum()

    def gaussian_blur2d(inp, kernel_size, sigma, border_type="reflect", separable=True):
        ky, kx = kernel_size
        sy, sx = sigma
        gy = gaussian_kernel1d(ky, sy, inp.device, inp.dtype)
        gx = gaussian_kernel1d(kx, sx, inp.device, inp.dtype)
        b, c, h, w = inp.shape
        x = F.pad(inp, (kx // 2, kx // 2, ky // 2, ky // 2), mode=border_type)
        x = F.conv2d(x, gx.view(1, 1, 1, kx).expand(c, 1, 1, kx), groups=c)
        x = F.conv2d(x, gy.view(1, 1, ky, 1).expand(c, 1, ky, 1), groups=c)
        return x

    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=None):
        xs = torch.linspace(0, width - 1, width, device=device, dtype=torch.float32)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=torch.float32)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        g = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)
        g = g.permute(1, 0, 2).unsqueeze(0)
        return g.to(dtype) if dtype is not None else g

    k = _stub("kornia")
    geo = _stub("kornia.geometry")
    tr = _stub("kornia.geometry.transform", remap=remap)
    fl = _stub("kornia.filters", gaussian_blur2d=gaussian_blur2d)
    ut = _stub("kornia.utils", create_meshgrid=create_meshgrid)
    k.geometry, geo.transform, k.filters, k.utils = geo, tr, fl, ut


def _xformers():
    def memory_efficient_attention(query, key, value, attn_bias=None, p=0.0, scale=None, op=None):
        return F.scaled_dot_product_attention(query, key, value, attn_mask=attn_bias, scale=scale)

    x = _stub("xformers")
    ops = _stub("xformers.ops", memory_efficient_attention=memory_efficient_attention)
    x.ops = ops


def install():
    """Make ``import animatediff / src / diffusers`` resolve to the reference tree."""
    global _done
    if _done:
        return
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import huggingface_hub

    for name in ("HfFolder", "cached_download", "whoami"):
        if not hasattr(huggingface_hub, name):
            setattr(huggingface_hub, name, type(name, (), {"get_token": staticmethod(lambda: None)})
                    if name == "HfFolder" else (lambda *a, **k: None))
    import huggingface_hub.constants as hc

    if not hasattr(hc, "hf_cache_home"):
        hc.hf_cache_home = "/tmp/hf_cache"
    _kornia()
    _xformers()
    fs = _stub("fairscale")
    fsn = _stub("fairscale.nn")
    fsc = _stub("fairscale.nn.checkpoint", checkpoint_wrapper=lambda m, *a, **k: m)
    fs.nn, fsn.checkpoint = fsn, fsc
    for name in ("imageio", "decord", "geocalib", "omegaconf", "loguru_stub"):
        if name not in sys.modules and importlib.util.find_spec(name) is None:
            _stub(name)
    if "omegaconf" in sys.modules and not hasattr(sys.modules["omegaconf"], "OmegaConf"):
        sys.modules["omegaconf"].OmegaConf = object

    # the vendored diffusers predates transformers 5: hide optional back-ends while it imports
    real_find_spec = importlib.util.find_spec
    hidden = {"transformers", "peft", "flax", "onnxruntime", "k_diffusion", "librosa", "note_seq", "accelerate",
              "xformers_real", "torchsde", "invisible_watermark", "compel", "ftfy", "bs4", "wandb", "omegaconf_real"}

    def find_spec(name, *a, **k):
        if name.split(".")[0] in hidden:
            return None
        return real_find_spec(name, *a, **k)

    importlib.util.find_spec = find_spec
    try:
        import diffusers  # noqa: F401  (the vendored tree)
        assert diffusers.__file__.startswith(REF), diffusers.__file__
    finally:
        importlib.util.find_spec = real_find_spec
    import diffusers.models.attention_processor as _ap

    _ap.xformers = sys.modules["xformers"]   # module-level `xformers = None` when the wheel is absent
    import src.modules.utils as smu

    smu.flush = lambda: None
    import src.models.MVGenModel as MV

    MV.flush = lambda: None
    if not torch.cuda.is_available():
        torch.cuda.empty_cache = lambda: None
    _done = True


SD21_UNET = dict(
    sample_size=96, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1,
    act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20),
    dual_cross_attention=False, use_linear_projection=True, upcast_attention=True,
)

UNET_ADDITIONAL = dict(  # configs/prompt-dual.yaml:16-45
    use_motion_module=True, use_inflated_groupnorm=True, motion_module_resolutions=(1, 2, 4, 8),
    motion_module_mid_block=True, motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=64,
                              temporal_attention_dim_div=1, zero_initialize=True),
    unet_use_cross_frame_attention=False, unet_use_temporal_attention=False, use_fps_condition=True,
    use_relative_postions="WithAdapter", use_ip_plus_cross_attention=True, ip_plus_condition="video",
    num_tokens=64, use_adapter_temporal_projection=True, compress_video_features=True, image_hidden_size=256,
    use_outpaint=True,
)


def tiny_unet_kwargs(c0=32, heads=(1, 2, 2, 2), cross_dim=64, img_hidden=16, num_tokens=8, mm_heads=2):
    """A structurally identical but tiny UNet3DConditionModel configuration for fixtures."""
    kw = dict(SD21_UNET)
    kw.update(UNET_ADDITIONAL)
    kw["block_out_channels"] = (c0, 2 * c0, 4 * c0, 4 * c0)
    kw["attention_head_dim"] = heads
    kw["cross_attention_dim"] = cross_dim
    kw["adapter_cross_attention_dim"] = cross_dim
    kw["image_cross_attention_dim"] = cross_dim
    kw["image_hidden_size"] = img_hidden
    kw["num_tokens"] = num_tokens
    kw["norm_num_groups"] = 32
    mk = dict(kw["motion_module_kwargs"])
    mk["num_attention_heads"] = mm_heads
    kw["motion_module_kwargs"] = mk
    return kw


def rerandomize_zero_init(module, seed=0, std=0.02):
    """Zero-initialised layers make random-init parity vacuous (SURVEY.md trap 4)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.numel() > 0 and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * std)


if __name__ == "__main__":
    install()
    from animatediff.models.unet import UNet3DConditionModel  # noqa
    from src.models.MVGenModel import MultiViewBaseModel  # noqa
    from diffusers import AutoencoderKL, DDIMScheduler  # noqa

    print("reference imported OK")
