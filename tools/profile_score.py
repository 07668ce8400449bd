"""One batched scoring call (16 views of 800x800) between cudaProfilerStart/Stop -- for ncu:
    ncu --profile-from-start off --set full --import-source on -k regex:'sort_downsweep|score_prologue' ... python tools/profile_score.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import metrics, synthetic

dev = torch.device("cuda:0")
b, h, w = int(os.environ.get("UB_PROFILE_VIEWS", 16)), 800, 800
imgs = [synthetic.scoring_image(h, w, seed=i, device=dev) for i in range(b)]
pred = torch.stack([i[0] for i in imgs]); std = torch.stack([i[1] for i in imgs]); gt = torch.stack([i[2] for i in imgs])
for _ in range(3):
    metrics.score_rgb_batch(pred, gt, std)
torch.cuda.synchronize()
torch.cuda.profiler.start()
metrics.score_rgb_batch(pred, gt, std)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
