"""How long the reduce / prologue / select calls of one view take on a side stream while the persistent compositing
kernels of the next view own the SMs (development aid).   python tools/diag_corun.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from uncertainty_nerf_gs_b200 import metrics, ops, pipeline, synthetic
from uncertainty_nerf_gs_b200.models import outputs as mo

dev = torch.device("cuda:0")
H, W, S, M = 840, 1297, 48, 5
R = H * W
members = [synthetic.ray_samples(R, S, seed=i, device=dev) for i in range(M)]
_, _, gt = synthetic.scoring_image(H, W, seed=0, device=dev)
outs = pipeline.render_members(members, H, W, 1 << 15)
red = mo.ensemble_reduce(outs)
n = H * W
z = metrics._z_table(dev)
cuts = metrics._tiled_cuts(n, 1)
side = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)


def post(evs):
    evs[0].record(side)
    r = mo.ensemble_reduce(outs)
    evs[1].record(side)
    pro = ops.score_prologue(r["rgb"].reshape(-1, 3), gt.reshape(-1, 3), r["rgb_std"].reshape(-1), [n], z,
                             nll_min_std=3e-2, sigma_from_var=True, want_vectors=True, want_coarse=True)
    evs[2].record(side)
    vec = pro["vectors"]
    ops.cut_select_sums([(vec[0], vec[1], vec[2]), (vec[1], vec[1], None), (vec[2], vec[2], None)], [n], cuts,
                        coarse=pro["coarse"])
    evs[3].record(side)


for busy in (False, True):
    for rep in range(3):
        torch.cuda.synchronize()
        evs = [ev() for _ in range(4)]
        if busy:
            for _ in range(3):
                pipeline.render_members(members, H, W, 1 << 15)
        with torch.cuda.stream(side):
            post(evs)
        torch.cuda.synchronize()
        print(f"compositing on the main stream: {busy}   reduce {evs[0].elapsed_time(evs[1]):.3f}  prologue "
              f"{evs[1].elapsed_time(evs[2]):.3f}  select {evs[2].elapsed_time(evs[3]):.3f} ms")
