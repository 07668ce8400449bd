"""Where does the host block when scoring runs on a side stream?  Development aid."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import metrics, pipeline, synthetic
from uncertainty_nerf_gs_b200.models import outputs as mo

dev = torch.device("cuda:0")
H, W, S, M = 840, 1297, 48, 5
members = [synthetic.ray_samples(H * W, S, seed=i, device=dev) for i in range(M)]
_, _, gt = synthetic.scoring_image(H, W, seed=0, device=dev)
side = torch.cuda.Stream()
for overlap in (False, True, False, True):
    acc = {"render": 0.0, "reduce": 0.0, "score": 0.0, "finish": 0.0}
    pend = None
    N = 20
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for it in range(N + 5):
        if it == 5:
            torch.cuda.synchronize()
            t_all = time.perf_counter()
            acc = {k: 0.0 for k in acc}
        t0 = time.perf_counter()
        outs = pipeline.render_members(members, H, W, 1 << 15)
        t1 = time.perf_counter()
        red = mo.ensemble_reduce(outs)
        t2 = time.perf_counter()
        if overlap:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                nxt = metrics.score_rgb_batch_async(red["rgb"], gt, red["rgb_std"])
            nxt.keep_alive = red
        else:
            nxt = metrics.score_rgb_batch_async(red["rgb"], gt, red["rgb_std"])
        t3 = time.perf_counter()
        if pend is not None:
            pend.finish()
        t4 = time.perf_counter()
        pend = nxt
        acc["render"] += t1 - t0; acc["reduce"] += t2 - t1; acc["score"] += t3 - t2; acc["finish"] += t4 - t3
    pend.finish()
    torch.cuda.synchronize()
    total = (time.perf_counter() - t_all) / N * 1e3
    print(json.dumps({"overlap": overlap, "ms_per_view": round(total, 3), **{k: round(v / N * 1e3, 3) for k, v in acc.items()}}))
