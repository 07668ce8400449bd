"""Where one bench step (configs[1]: 5 members -> composite -> reduce -> score) spends its time: device time per phase
(CUDA events) and host enqueue / tail time (perf_counter).  Development aid."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import metrics, pipeline, synthetic
from uncertainty_nerf_gs_b200.models import outputs as mo

dev = torch.device("cuda:0")
H, W, S, M = 840, 1297, 48, 5
R = H * W
members = [synthetic.ray_samples(R, S, seed=i, device=dev) for i in range(M)]
_, _, gt = synthetic.scoring_image(H, W, seed=0, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
N = 20
acc = {"composite": 0.0, "reduce": 0.0, "score": 0.0, "host_enqueue": 0.0, "host_tail": 0.0, "step_wall": 0.0}
for it in range(N + 3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e = [ev() for _ in range(4)]
    e[0].record()
    outs = pipeline.render_members(members, H, W, 1 << 15)
    e[1].record()
    red = mo.ensemble_reduce(outs)
    e[2].record()
    pend = metrics.score_rgb_batch_async(red["rgb"], gt, red["rgb_std"])
    e[3].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    pend.finish()
    t3 = time.perf_counter()
    if it >= 3:
        acc["composite"] += e[0].elapsed_time(e[1]); acc["reduce"] += e[1].elapsed_time(e[2]); acc["score"] += e[2].elapsed_time(e[3])
        acc["host_enqueue"] += (t1 - t0) * 1e3; acc["host_tail"] += (t3 - t2) * 1e3; acc["step_wall"] += (t3 - t0) * 1e3
print(json.dumps({k: round(v / N, 4) for k, v in acc.items()}))

# host cost of the individual wrappers (no sync in between; the device queue never drains: pure host time)
import cProfile, pstats, io
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(10):
    outs = pipeline.render_members(members, H, W, 1 << 15)
    red = mo.ensemble_reduce(outs)
    pend = metrics.score_rgb_batch_async(red["rgb"], gt, red["rgb_std"])
    pend.finish()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
