"""Locate torch-side copy kernels inside one denoise step (they are not ours): torch.profiler with python stacks."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

size = bench.SIZES["c3"]
pipe, inp = bench.build_native("cuda:0", size)
lat = [inp["pano_latent"], inp["pers_latent"]]
def one_step(i):
    lat[0], lat[1] = pipe.denoise(lat[0], lat[1], inp["pano_mask"], inp["pers_masks"], inp["pano_masked"], inp["pers_masked"],
                                  inp["cond"], inp["cameras"], 50, 7.5, step_range=(i, i + 1))
random.seed(0)
one_step(0); one_step(1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    one_step(2)
    torch.cuda.synchronize()
avg = prof.key_averages(group_by_stack_n=12)
rows = []
for e in avg:
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if t > 50 and e.key.startswith("aten::"):
        st = [x for x in (e.stack or []) if "imagine360" in x or "bench.py" in x]
        rows.append((t, e.key, e.count, st[:4]))
rows.sort(reverse=True, key=lambda r: r[0])
for r in rows[:30]:
    print(f"{r[0]:10.1f} us  x{r[2]:<4d} {r[1]:24s} {r[3]}")
