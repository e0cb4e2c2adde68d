"""Deterministic synthetic parameters / inputs shared by tools/make_golden.py (which runs the unmodified
reference in the build container) and the tests (which run the oracle and the CUDA path anywhere).

Weights are never stored in fixtures: a fixture records {key: shape} and a seed, and both sides
regenerate identical tensors with a CPU torch.Generator (mt19937 + fixed transform: reproducible
for a given torch build, and the GPU box runs the same image)."""
from __future__ import annotations

import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BUFFER_SUFFIXES = ("pos_encoder.pe", "pe.freq_bands")


def synth_state(shapes: dict, seed: int, dtype=torch.float32) -> dict:
    """{key: shape} -> {key: tensor}; norm scales ~ 1 + 0.1 N, biases ~ 0.1 N, matrices ~ N / sqrt(fan_in)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        if key.endswith(BUFFER_SUFFIXES):
            continue
        t = torch.randn(shape, generator=g)
        if len(shape) <= 1:
            if key.endswith(".weight"):
                t = 1.0 + 0.1 * t
            else:
                t = 0.1 * t
        elif key.endswith("latents"):
            t = t / shape[-1] ** 0.5
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = t / fan_in ** 0.5
        sd[key] = t.to(dtype)
    return sd


def synth_tensor(shape, seed: int, scale: float = 1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(tuple(shape), generator=g) * scale).to(dtype)


def load(name: str):
    return torch.load(os.path.join(GOLDEN_DIR, name), map_location="cpu", weights_only=False)


def tiny_cameras(m: int):
    """A small, fixed camera set (FoV 90) used by the tiny fixtures."""
    thetas = [-150.0, -30.0, 95.0, 180.0, 20.0][:m]
    phis = [35.0, -20.0, 0.0, 60.0, -52.0][:m]
    return dict(FoV=[90] * m, theta=thetas, phi=phis)
