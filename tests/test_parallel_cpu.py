"""world_size-2 gloo test of the clip-sharding plumbing (no GPU): every clip is processed exactly once, timings reduce
with MAX, results gather on rank 0, and the per-clip seeds are distinct."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from imagine360_b200.host import parallel as P
    mine = P.clips_for_rank(5, rank, world)
    local = {c: torch.full((2,), float(P.seeds_for_clip(996995, c))) for c in mine}
    dist.barrier()
    t = P.max_over_ranks(1.0 + rank)
    got = P.gather_results(local)
    if rank == 0:
        merged = {}
        for d in got:
            merged.update(d)
        q.put((t, sorted(merged), [float(merged[c][0]) for c in sorted(merged)]))
    dist.destroy_process_group()


def test_two_rank_clip_map():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, clips, seeds = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert t == 2.0
    assert clips == [0, 1, 2, 3, 4]
    assert seeds == [996995.0 + c for c in range(5)]


def test_dropin_install_registers_boundary():
    import imagine360_b200.dropin as d
    done = d.install()
    from animatediff.pipelines.pipeline_animation_inference_dual import AnimationPipeline
    from src.models.MVGenModel import MultiViewBaseModel
    from diffusers import AutoencoderKL, DDIMScheduler
    from imagine360_b200.host import mvgen, pipeline, vae, ddim
    assert AnimationPipeline is pipeline.AnimationPipeline and MultiViewBaseModel is mvgen.MultiViewBaseModel
    assert AutoencoderKL is vae.AutoencoderKL and DDIMScheduler is ddim.DDIMScheduler
    from animatediff.utils.video_mask import get_anchor_target
    from src.utils.pano_utils.Equirec2Perspec import Equirectangular
    from imagine360_b200.host import preprocess
    assert get_anchor_target is preprocess.get_anchor_target and Equirectangular is preprocess.Equirectangular
    assert set(done) == set(d.BOUNDARY)
    for name in list(sys.modules):
        if name.split(".")[0] in ("animatediff", "src", "diffusers") and not getattr(sys.modules[name], "__file__", None):
            del sys.modules[name]
