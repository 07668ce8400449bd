"""Timeline of the graphed step (development aid): per view, when its compositing graph and its reduce + scoring graph
start and end on the device, relative to the first view -- shows whether the scoring of view i still runs when the
compositing of view i + 2 (same buffer set) wants to start.   python tools/diag_overlap.py [views]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import pipeline, synthetic

dev = torch.device("cuda:0")
H, W, S, M = 840, 1297, 48, 5
R = H * W
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
members = [synthetic.ray_samples(R, S, seed=i, device=dev) for i in range(M)]
_, _, gt = synthetic.scoring_image(H, W, seed=0, device=dev)
g = pipeline.GraphedViews(H, W, 1 << 15)
for _ in range(4):
    g.launch(members, gt).finish()
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
marks = []
side = g._side
main = torch.cuda.current_stream(dev)
orig_replay = {}
pend = []
origin = ev(); origin.record(main)
for v in range(N):
    turn = g._count & 1
    slot = g._slots[(g._key(members), turn)]
    c0, c1, p0, p1 = ev(), ev(), ev(), ev()
    # same sequence as GraphedViews.launch, with timing events
    if slot.pending is not None:
        slot.pending.finish()
    if g._last_post[turn] is not None:
        main.wait_event(g._last_post[turn])
    c0.record(main); slot.comp.replay(); c1.record(main)
    slot.comp_done.record(main)
    with torch.cuda.stream(side):
        side.wait_event(slot.comp_done)
        slot.gt.copy_(gt, non_blocking=True)
        p0.record(side); slot.post.replay(); p1.record(side)
        slot.post_done.record(side)
    g._last_post[turn] = slot.post_done
    g._count += 1
    from uncertainty_nerf_gs_b200 import metrics
    slot.pending = metrics.PendingScores(slot.packed_host, slot.packed_dev, slot.post_done, slot.b, slot.n, slot.c, slot.cuts_one)
    marks.append((c0, c1, p0, p1))
torch.cuda.synchronize()
for v, (c0, c1, p0, p1) in enumerate(marks):
    print(f"view {v:2d}: comp {origin.elapsed_time(c0):7.3f} -> {origin.elapsed_time(c1):7.3f} ms ({c0.elapsed_time(c1):.3f})   "
          f"post {origin.elapsed_time(p0):7.3f} -> {origin.elapsed_time(p1):7.3f} ms ({p0.elapsed_time(p1):.3f})")
