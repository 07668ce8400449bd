"""Why is a view of the 64-view sweep slower than the repeated view of the headline?  Times the batched compositing
call (CUDA events) while rotating over P resident member sets, P = 1, 2, 4, 8 (8.4 GB each)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import pipeline, synthetic
from uncertainty_nerf_gs_b200.build import build_library

build_library()
dev = torch.device("cuda:0")
H, W, S, M = 840, 1297, 48, 5
sets = []
for P in (1, 2, 4, 8):
    while len(sets) < P:
        sets.append([synthetic.ray_samples(H * W, S, seed=100 * len(sets) + i, device=dev) for i in range(M)])
    torch.cuda.synchronize()
    for _ in range(3):
        pipeline.render_members(sets[0], H, W, 1 << 15)
    timers = []
    for i in range(24):
        pipeline.render_members(sets[i % P], H, W, 1 << 15, timers)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) / c for a, b, c in timers]
    print(json.dumps({"resident_sets": P, "GB": P * M * H * W * 1536 / 1e9, "ms_per_member_launch_mean": sum(ms) / len(ms),
                      "min": min(ms), "max": max(ms), "mem_allocated_GB": torch.cuda.memory_allocated() / 1e9}), flush=True)
# same with the graphed evaluator end to end (views/s)
gts = [torch.rand(H, W, 3, device=dev) for _ in range(8)]
for P in (1, 8):
    gv = pipeline.GraphedViews(H, W, 1 << 15)
    for i in range(2 * P + 2):
        gv.launch(sets[i % P], gts[i % 8]).finish()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    prev = None
    for i in range(32):
        cur = gv.launch(sets[i % P], gts[i % 8])
        if prev is not None:
            prev.finish()
        prev = cur
    prev.finish()
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"graphed_views_rotating_over": P, "ms_per_view": a.elapsed_time(b) / 32}), flush=True)
