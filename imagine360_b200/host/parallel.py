"""Outer data-parallel map over independent clips.

The reference is single-process, batch 1 (pipeline...dual.py:598,:672); one clip (all frames, both branches, all steps)
is the unit of work and nothing is exchanged inside a step.  Each rank holds a full replica and takes clips
``rank, rank + world, ...``; the only collectives are a barrier around the timed region, a MAX all-reduce of the elapsed
time and an optional gather of the finished videos."""
from __future__ import annotations

import torch
import torch.distributed as dist


def clips_for_rank(n_clips: int, rank: int, world: int) -> list[int]:
    return list(range(rank, n_clips, world))


def seeds_for_clip(base_seed: int, clip: int) -> int:
    """inference_dual_p2e.py:349-351 seeds torch / random / numpy with one global seed; clip i uses seed + i."""
    return base_seed + clip


def max_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local: dict) -> list[dict] | None:
    """{clip index: tensor} of every rank collected on rank 0 (None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    return out
