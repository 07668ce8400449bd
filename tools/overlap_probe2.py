import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import metrics, pipeline, synthetic
dev = torch.device("cuda:0")
H, W, S, M = 840, 1297, 48, 5
members = [synthetic.ray_samples(H * W, S, seed=i, device=dev) for i in range(M)]
_, _, gt = synthetic.scoring_image(H, W, seed=0, device=dev)
def loop(overlap, use_timers, n=30):
    timers = [] if use_timers else None
    pend = None
    for _ in range(5):
        pipeline.evaluate_view_async(members, gt, H, W, 1 << 15, overlap_scoring=overlap).finish()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enq = fin = 0.0
    e0.record()
    for _ in range(n):
        a = time.perf_counter()
        nxt = pipeline.evaluate_view_async(members, gt, H, W, 1 << 15, timers=timers, overlap_scoring=overlap)
        b = time.perf_counter()
        if pend is not None:
            pipeline.pack_record(0, pend.finish())
        c = time.perf_counter()
        enq += b - a; fin += c - b
        pend = nxt
    pend.finish()
    e1.record(); torch.cuda.synchronize()
    comp = sum(x.elapsed_time(y) for x, y in timers) / len(timers) if timers else None
    print(json.dumps({"overlap": overlap, "timers": use_timers, "ms_per_step": round(e0.elapsed_time(e1) / n, 4),
                      "enqueue": round(enq / n * 1e3, 3), "wait_tail": round(fin / n * 1e3, 3), "composite_ms": comp}))
for ov in (False, True):
    for tm in (False, True):
        loop(ov, tm)
