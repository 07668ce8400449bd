"""Host-side cost of the wrappers (time to *enqueue*, device work is tiny: 4096-ray batches).  Development aid."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import metrics, ops, pipeline, synthetic
from uncertainty_nerf_gs_b200.models import outputs as mo

dev = torch.device("cuda:0")
def bench(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    dt = (time.perf_counter() - t) / n
    torch.cuda.synchronize()
    return dt * 1e6
m = synthetic.ray_samples(4096, 48, seed=0, device=dev)
res = {}
res["composite_rays"] = bench(lambda: ops.composite_rays(m["density"], m["deltas"], m["starts"], m["ends"], m["rgb"], m["beta"], rays_per_chunk=1 << 15))
res["active_nerfacto_outputs"] = bench(lambda: mo.active_nerfacto_outputs(m["density"], m["deltas"], m["starts"], m["ends"], m["rgb"], m["beta"], rays_per_chunk=1 << 15))
h, w = 64, 64
outs = [{k: v.view(h, w, -1) for k, v in mo.active_nerfacto_outputs(m["density"], m["deltas"], m["starts"], m["ends"], m["rgb"], m["beta"]).items() if k != "density"} for _ in range(5)]
res["ensemble_reduce"] = bench(lambda: mo.ensemble_reduce(outs))
red = mo.ensemble_reduce(outs)
_, _, gt = synthetic.scoring_image(h, w, seed=0, device=dev)
res["score_rgb_batch_async"] = bench(lambda: metrics.score_rgb_batch_async(red["rgb"], gt, red["rgb_std"]))
z = metrics._z_table(dev)
res["score_prologue"] = bench(lambda: ops.score_prologue(red["rgb"].reshape(-1, 3), gt.reshape(-1, 3), red["rgb_std"].reshape(-1), [h * w], z, 0.03))
v = torch.rand(3 * h * w, device=dev)
res["segmented_sort"] = bench(lambda: ops.segmented_sort(v, [h * w] * 3, want_perm=True, want_keys=True))
res["torch.empty"] = bench(lambda: torch.empty(100, device=dev))
res["event_record"] = bench(lambda: torch.cuda.Event().record())
res["_stream"] = bench(lambda: ops._stream())
print(json.dumps({k: round(v, 1) for k, v in res.items()}))

# ---- whole-view enqueue cost with and without the scoring side stream (tiny device work) ----
hh, ww = 64, 64
members = [synthetic.ray_samples(hh * ww, 48, seed=i, device=dev) for i in range(5)]
for overlap in (False, True):
    pend = []
    def run():
        pend.append(pipeline.evaluate_view_async(members, gt, hh, ww, 1 << 15, overlap_scoring=overlap))
        if len(pend) > 1:
            pend.pop(0).finish()
    us = bench(run, n=100)
    print(json.dumps({"evaluate_view_async+finish host us": round(us, 1), "overlap_scoring": overlap}))
    for p_ in pend: p_.finish()
    pend.clear()
side = torch.cuda.Stream()
def sw():
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        pass
print(json.dumps({"stream switch us": round(bench(sw), 1)}))

# ---- where the host time of one scoring call goes (800x800) ----
import cProfile, pstats, io
p8, s8, g8 = synthetic.scoring_image(800, 800, seed=0, device=dev)
for _ in range(5):
    metrics.score_rgb_batch_async(p8, g8, s8).finish()
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
pend = [metrics.score_rgb_batch_async(p8, g8, s8) for _ in range(50)]
pr.disable()
[x.finish() for x in pend]
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4500])
