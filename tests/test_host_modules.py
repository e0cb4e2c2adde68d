"""CPU checks of the host-side mirror: parameter names/shapes are exactly the reference's (taken from the fixtures
that tools/make_golden.py recorded from the unmodified reference modules), and the C-ABI library exports every
symbol the header declares."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_util import load  # noqa: E402

TINY_KW = dict(
    sample_size=96, in_channels=4, out_channels=4, flip_sin_to_cos=True, freq_shift=0,
    block_out_channels=(32, 64, 128, 128), layers_per_block=2, norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=32,
    attention_head_dim=(1, 2, 4, 4), use_linear_projection=True, upcast_attention=True,
    use_motion_module=True, use_inflated_groupnorm=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=True,
    motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=2, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=64,
                              temporal_attention_dim_div=1, zero_initialize=True),
    unet_use_cross_frame_attention=False, unet_use_temporal_attention=False, use_fps_condition=True,
    use_relative_postions="WithAdapter", use_ip_plus_cross_attention=True, ip_plus_condition="video", num_tokens=16,
    use_adapter_temporal_projection=True, compress_video_features=True, image_hidden_size=8, use_outpaint=True,
    adapter_cross_attention_dim=32, image_cross_attention_dim=32)


def tiny_unet():
    from imagine360_b200.host.unet3d import UNet3DConditionModel
    return UNet3DConditionModel(**TINY_KW)


def shapes(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def test_unet_state_dict_matches_reference():
    assert shapes(tiny_unet()) == load("unet3d.pt")["shapes"]


def test_mvgen_state_dict_matches_reference():
    from imagine360_b200.host.mvgen import MultiViewBaseModel
    mv = MultiViewBaseModel(tiny_unet(), tiny_unet())
    assert shapes(mv) == load("mvgen.pt")["shapes"]


def test_full_size_unet_parameter_count():
    """SD-2.1 topology + configs/prompt-dual.yaml kwargs -> 1 424.2 M parameters."""
    from imagine360_b200.host.unet3d import UNet3DConditionModel
    from imagine360_b200.host.config import FULL_UNET_KWARGS
    with torch.device("meta"):
        m = UNet3DConditionModel(**FULL_UNET_KWARGS)
    n = sum(p.numel() for p in m.parameters())
    assert abs(n / 1e6 - 1424.2) < 0.1, n


def test_library_exports_header_symbols():
    from imagine360_b200 import _lib
    L = _lib.lib()
    names = _lib.exported_symbols()
    assert len(names) >= 14
    for s in names:
        assert hasattr(L, s), s


def test_ddim_schedule_matches_reference():
    from imagine360_b200.host.ddim import DDIMScheduler
    g = load("ddim.pt")
    s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear", steps_offset=1,
                      clip_sample=False, prediction_type="v_prediction", rescale_betas_zero_snr=True)
    assert torch.allclose(s.alphas_cumprod, g["alphas_cumprod"], rtol=1e-6, atol=0)
    for n in (50, 25):
        s.set_timesteps(n)
        assert torch.equal(s.timesteps, g[f"timesteps_{n}"])


def test_fused_cross_attention_limits_are_host_checkable():
    """the host picks the fused text + image-prompt kernel only inside its limits (head_dim 64, <= 96 keys per branch,
    at most three 64-key P blocks); everything else takes the generic two-call path"""
    from imagine360_b200 import ops
    assert ops.cross_attention_text_ip_supported(64, 77, 64)          # the production shape
    assert ops.cross_attention_text_ip_supported(64, 77, 16) and ops.cross_attention_text_ip_supported(64, 1, 1)
    assert not ops.cross_attention_text_ip_supported(32, 77, 64)      # WarpAttn's head_dim
    assert not ops.cross_attention_text_ip_supported(64, 128, 64)     # > 96 keys in one branch
    assert not ops.cross_attention_text_ip_supported(64, 96, 96)      # 2 + 2 P blocks
