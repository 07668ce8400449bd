"""Time the AUSE cut sums through the select path and through the sort path (CUDA events, device already warm).

python tools/perf_select.py [views] [H] [W]   -> JSON lines
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uncertainty_nerf_gs_b200 import metrics as M, ops  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    views = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    w = int(sys.argv[3]) if len(sys.argv) > 3 else 800
    floor = float(os.environ.get("UB_PERF_STD_FLOOR", "0.03"))
    n = h * w
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.rand(views, h, w, 3, device="cuda", generator=g)
    std = torch.clamp(0.1 * torch.rand(views, h, w, 1, device="cuda", generator=g), min=floor)
    gt = torch.clamp(pred + std * torch.randn(views, h, w, 3, device="cuda", generator=g), 0, 1)
    lens = [n] * views
    z = M._z_table(pred.device)
    pro = ops.score_prologue(pred.reshape(-1, 3), gt.reshape(-1, 3), std.reshape(-1), lens, z, nll_min_std=3e-2,
                             sigma_from_var=True, want_vectors=True)
    vec = pro["vectors"]
    cuts = np.tile(M.ause_cut_counts(n)[None, :], (views, 1))
    out = {"views": views, "h": h, "w": w, "std_floor": floor}
    os.environ["UB_AUSE_SORT"] = "0"
    out["ms_select_sums"] = timed(lambda: M._ause_sums(vec, lens, cuts))
    out["ms_score_select"] = timed(lambda: M.score_rgb_batch_async(pred, gt, std))
    a = M._ause_sums(vec, lens, cuts)
    os.environ["UB_AUSE_SORT"] = "1"
    out["ms_sort_sums"] = timed(lambda: M._ause_sums(vec, lens, cuts))
    out["ms_score_sort"] = timed(lambda: M.score_rgb_batch_async(pred, gt, std))
    b = M._ause_sums(vec, lens, cuts)
    out["max_rel_dev"] = float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())
    out["images_s_select"] = views / out["ms_score_select"] * 1e3
    out["images_s_sort"] = views / out["ms_score_sort"] * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
