"""Host-side mirror of ``diffusers.DDIMScheduler`` as configured by configs/prompt-dual.yaml:48-56.

Same constructor kwargs / attributes / ``set_timesteps`` / ``step(...).prev_sample`` surface as the reference
(diffusers/schedulers/scheduling_ddim.py:113-372).  The schedule lives on the host as python floats (the reference
indexes a CPU ``alphas_cumprod`` with a CUDA timestep every step -> one device sync per step; here the timesteps are a
host list), and the update itself is the fused CFG + DDIM kernel when driven by the pipeline loop.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from .. import ops


def _rescale_zero_terminal_snr(betas):
    """scheduling_ddim.py:77-111"""
    alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    a0, aT = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt = (alphas_bar_sqrt - aT) * (a0 / (a0 - aT))
    alphas_bar = alphas_bar_sqrt ** 2
    return 1 - torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 rescale_betas_zero_snr=False, **kwargs):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        if rescale_betas_zero_snr:
            betas = _rescale_zero_terminal_snr(betas)
        if prediction_type != "v_prediction" or clip_sample:
            raise NotImplementedError("the Imagine360 path runs v_prediction without sample clipping (yaml:53-55)")
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                      clip_sample=clip_sample, prediction_type=prediction_type, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule,
                                      rescale_betas_zero_snr=rescale_betas_zero_snr)
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps_host = [int(t) for t in ts]
        self.timesteps = torch.from_numpy(ts).to(device)

    def coefficients(self, t: int):
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)) -- fp32-rounded like the reference's 0-dim tensors."""
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return tuple(float(v) for v in (a_t ** 0.5, (1 - a_t) ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5))

    def step(self, model_output, timestep, sample, eta: float = 0.0, **kwargs):
        """v-prediction, eta = 0 (scheduling_ddim.py:251-372).  Runs the fused kernel with guidance disabled
        (uncond == cond == model_output), which reduces to the plain DDIM update with identical rounding."""
        if eta != 0.0:
            raise NotImplementedError("eta > 0")
        sa, sb, sap, sbp = self.coefficients(int(timestep))
        mo = model_output.to(torch.bfloat16).contiguous()
        prev = ops.cfg_ddim_step(sample.to(torch.bfloat16).contiguous(), mo, mo, 0.0, sa, sb, sap, sbp)
        return SimpleNamespace(prev_sample=prev.to(sample.dtype))
