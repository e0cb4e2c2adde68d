"""Model topology used by the reference run: public SD-2.1 ``unet/config.json`` / ``vae/config.json`` values merged with
``unet_additional_kwargs`` of configs/prompt-dual.yaml:16-45 (weights are not in the reference tree)."""

FULL_UNET_KWARGS = dict(
    sample_size=96, in_channels=4, out_channels=4, flip_sin_to_cos=True, freq_shift=0,
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, norm_num_groups=32, norm_eps=1e-5,
    cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20), use_linear_projection=True, upcast_attention=True,
    # configs/prompt-dual.yaml:16-45
    use_motion_module=True, use_inflated_groupnorm=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=True,
    motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
                              temporal_position_encoding=True, temporal_position_encoding_max_len=64,
                              temporal_attention_dim_div=1, zero_initialize=True),
    unet_use_cross_frame_attention=False, unet_use_temporal_attention=False, use_fps_condition=True,
    use_relative_postions="WithAdapter", use_ip_plus_cross_attention=True, ip_plus_condition="video", num_tokens=64,
    use_adapter_temporal_projection=True, compress_video_features=True, image_hidden_size=256, use_outpaint=True,
)

FULL_VAE_KWARGS = dict(
    in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
    block_out_channels=(128, 256, 512, 512), layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=32,
    sample_size=768)

SCHEDULER_KWARGS = dict(  # configs/prompt-dual.yaml:48-56
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear", steps_offset=1, clip_sample=False,
    prediction_type="v_prediction", rescale_betas_zero_snr=True)
