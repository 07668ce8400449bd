"""One batched scoring call (16 views of 800x800) between cudaProfilerStart/Stop -- for ncu:
    ncu --profile-from-start off --set full --import-source on -k regex:'sort_downsweep|score_prologue' ... python tools/profile_score.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import metrics, synthetic

dev = torch.device("cuda:0")
b = int(os.environ.get("UB_PROFILE_VIEWS", 16))
h, w = int(os.environ.get("UB_PROFILE_H", 800)), int(os.environ.get("UB_PROFILE_W", 800))
if os.environ.get("UB_PROFILE_BENCH_DATA"):   # the images bench.py scores: a composited render against an unrelated ground truth
    from uncertainty_nerf_gs_b200 import pipeline
    out = pipeline.render_members([synthetic.ray_samples(h * w, 48, seed=0, device=dev)], h, w, 1 << 15)[0]
    _, _, g1 = synthetic.scoring_image(h, w, seed=0, device=dev)
    pred, std, gt = (t[None].expand(b, *t.shape).contiguous() for t in (out["rgb"].clone(), out["rgb_std"].clone(), g1))
else:
    imgs = [synthetic.scoring_image(h, w, seed=i, device=dev) for i in range(b)]
    pred = torch.stack([i[0] for i in imgs]); std = torch.stack([i[1] for i in imgs]); gt = torch.stack([i[2] for i in imgs])
for _ in range(3):
    metrics.score_rgb_batch(pred, gt, std)
torch.cuda.synchronize()
torch.cuda.profiler.start()
metrics.score_rgb_batch(pred, gt, std)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
