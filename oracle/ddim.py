"""Oracle: DDIMScheduler as configured by configs/prompt-dual.yaml:48-56
(linear betas 0.00085..0.012, zero-terminal-SNR rescale, steps_offset 1, v-prediction, eta 0).
Follows diffusers/schedulers/scheduling_ddim.py:77-111 (rescale), :235-249 (set_timesteps), :251-372 (step)."""
from __future__ import annotations

import numpy as np
import torch


def rescale_zero_terminal_snr(betas):
    alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    a0, aT = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt = (alphas_bar_sqrt - aT) * (a0 / (a0 - aT))
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
    return 1 - alphas


class DDIM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 rescale_betas_zero_snr=True, set_alpha_to_one=True):
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        if rescale_betas_zero_snr:
            betas = rescale_zero_terminal_snr(betas)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.T = num_train_timesteps
        self.steps_offset = steps_offset

    def set_timesteps(self, n):
        self.n = n
        ratio = self.T // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.steps_offset
        return self.timesteps

    def coefficients(self, t: int):
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)) as fp32 0-dim tensors."""
        prev = t - self.T // self.n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return a_t ** 0.5, (1 - a_t) ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5

    def step(self, v, t: int, x):
        """v-prediction, eta = 0, no clipping.  Each product is (fp32 scalar) x (tensor in x.dtype) ->
        x.dtype, exactly as in the reference (0-dim CPU tensors do not promote)."""
        sa, sb, sap, sbp = self.coefficients(int(t))
        x0 = sa * x - sb * v
        eps = sa * v + sb * x
        return sap * x0 + sbp * eps


def cfg_combine(pred, scale=7.5):
    """pipeline_animation_inference_dual.py:789-795: uncond + s * (text - uncond)."""
    u, c = pred.chunk(2)
    return u + scale * (c - u)
