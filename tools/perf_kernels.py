"""Per-kernel device timings (CUDA events, rotating input sets larger than L2) -- development aid.
Usage: python tools/perf_kernels.py [composite|reduce|score|sort|laplace|splat ...]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uncertainty_nerf_gs_b200 import ops, synthetic, metrics
from uncertainty_nerf_gs_b200.build import build_library

build_library()
dev = torch.device("cuda:0")
H, W, S = 840, 1297, 48
R = H * W


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def composite():
    sets = [synthetic.ray_samples(R, S, seed=i, device=dev) for i in range(3)]
    def run(i):
        m = sets[i % 3]
        ops.composite_rays(m["density"], m["deltas"], m["starts"], m["ends"], m["rgb"], m["beta"], rays_per_chunk=1 << 15)
    ms = timeit(run)
    print(json.dumps({"kernel": "composite_rays", "ms": ms, "GBs": 1576 * R / ms / 1e6, "Grays_s": R / ms / 1e6}))
    for r in (4096, 32768):
        small = [{k: v[:r].contiguous() for k, v in s.items()} for s in sets]
        def run2(i):
            m = small[i % 3]
            ops.composite_rays(m["density"], m["deltas"], m["starts"], m["ends"], m["rgb"], m["beta"], rays_per_chunk=1 << 15)
        ms = timeit(run2, iters=50)
        print(json.dumps({"kernel": f"composite_rays[{r}]", "ms": ms, "GBs": 1576 * r / ms / 1e6}))


def backward():
    from uncertainty_nerf_gs_b200.autograd import composite_rays_train
    for r in (4096, R):
        m = synthetic.ray_samples(r, S, seed=1, device=dev)
        leaves = {k: m[k].clone().requires_grad_(True) for k in ("density", "rgb", "beta")}
        def run(i):
            out = composite_rays_train(leaves["density"], m["deltas"], m["starts"], m["ends"], leaves["rgb"], leaves["beta"])
            (out["rgb"].sum() + out["rgb_var"].sum() + out["accumulation"].sum()).backward()
            for v in leaves.values():
                v.grad = None
        ms = timeit(run, iters=10)
        fwd = timeit(lambda i: ops.composite_rays(m["density"], m["deltas"], m["starts"], m["ends"], m["rgb"], m["beta"],
                                                  eval_mode=False, return_weights=True), iters=10)
        print(json.dumps({"kernel": f"composite_rays train fwd+bwd [{r} rays]", "ms_fwd_bwd_with_torch_loss": ms, "ms_fwd_only": fwd,
                          "Mrays_s": r / ms / 1e3}))


def reduce():
    for K in (5, 10):
        ms_ = [torch.rand(R, 3, device=dev) for _ in range(K)]
        ds = [torch.rand(R, 1, device=dev) for _ in range(K)]
        t3 = timeit(lambda i: ops.reduce_members(ms_, "std"))
        t1 = timeit(lambda i: ops.reduce_members(ds, "std"))
        t0 = timeit(lambda i: ops.reduce_members(ds, None))
        print(json.dumps({"kernel": f"reduce K={K}", "ms_c3_std": t3, "GBs_c3": (K * 12 + 16) * R / t3 / 1e6,
                          "ms_c1_std": t1, "GBs_c1": (K * 4 + 8) * R / t1 / 1e6, "ms_c1_mean": t0,
                          "GBs_c1_mean": (K * 4 + 4) * R / t0 / 1e6}))


def score():
    for (h, w, b) in ((840, 1297, 1), (800, 800, 1), (800, 800, 16)):
        imgs = [synthetic.scoring_image(h, w, seed=i, device=dev) for i in range(b)]
        pred = torch.stack([i[0] for i in imgs]); std = torch.stack([i[1] for i in imgs]); gt = torch.stack([i[2] for i in imgs])
        ms = timeit(lambda i: metrics.score_rgb_batch(pred, gt, std), iters=10)
        pend = []
        def streamed(i):          # call i+1 is enqueued before call i's record is read back
            pend.append(metrics.score_rgb_batch_async(pred, gt, std))
            if len(pend) > 1:
                pend.pop(0).finish()
        ms_stream = timeit(streamed, iters=12)
        for q in pend:
            q.finish()
        n = h * w
        z = metrics._z_table(dev)
        tp = timeit(lambda i: ops.score_prologue(pred.reshape(-1, 3), gt.reshape(-1, 3), std.reshape(-1), [n] * b, z, 0.03))
        var = (std ** 2).reshape(-1)
        ts = timeit(lambda i: ops.segmented_sort(var, [n] * b, want_perm=True, want_keys=False))
        tk = timeit(lambda i: ops.segmented_sort(var, [n] * b, want_perm=False, want_keys=True))
        pro = ops.score_prologue(pred.reshape(-1, 3), gt.reshape(-1, 3), std.reshape(-1), [n] * b, z, 0.03, want_coarse=True)
        vec = pro["vectors"]
        tpc = timeit(lambda i: ops.score_prologue(pred.reshape(-1, 3), gt.reshape(-1, 3), std.reshape(-1), [n] * b, z, 0.03,
                                                  want_coarse=True))
        cuts = np.tile(metrics.ause_cut_counts(n)[None, :], (b, 1))
        prev = os.environ.get("UB_AUSE_SORT")
        os.environ["UB_AUSE_SORT"] = "0"
        t_sel = timeit(lambda i: metrics._ause_sums(vec, [n] * b, cuts))
        t_sel_c = timeit(lambda i: metrics._ause_sums(vec, [n] * b, cuts, pro["coarse"]))
        os.environ["UB_AUSE_SORT"] = "1"
        t_srt = timeit(lambda i: metrics._ause_sums(vec, [n] * b, cuts))
        if prev is None:
            os.environ.pop("UB_AUSE_SORT")
        else:
            os.environ["UB_AUSE_SORT"] = prev
        print(json.dumps({"kernel": f"ause cut sums {w}x{h} x{b}", "ms_select": t_sel, "ms_select_given_coarse_hist": t_sel_c,
                          "ms_prologue_with_coarse_hist": tpc, "ms_sort_and_cut_sums": t_srt,
                          "Mkeys_s_select": 3 * n * b / t_sel / 1e3}))
        print(json.dumps({"kernel": f"score {w}x{h} x{b}", "ms_total": ms, "images_s": b / ms * 1e3, "ms_streamed": ms_stream, "images_s_streamed": b / ms_stream * 1e3, "ms_prologue": tp,
                          "prologue_GBs": 40 * n * b / tp / 1e6, "ms_sort_pairs": ts, "ms_sort_keys": tk,
                          "sort_pairs_Mkeys_s": n * b / ts / 1e3}))


def laplace():
    P = 1 << 20
    lap = synthetic.laplace_head(P, 64, 3, 100, seed=0, device=dev)
    theta = lap["mu_q"][None] + lap["eps_draws"] / torch.sqrt(lap["ggn"] + 1.0)[None]
    ms = timeit(lambda i: ops.laplace_ll_moments(lap["x"], theta, 3, "sigmoid"), iters=5, warm=1)
    print(json.dumps({"kernel": "laplace rgb head", "ms": ms, "Mpoints_s": P / ms / 1e3, "TFLOPs": 2 * 64 * 3 * 100 * P / ms / 1e9}))
    th1 = theta[:, :65].contiguous()
    ms = timeit(lambda i: ops.laplace_ll_moments(lap["x"], th1, 1, "exp"), iters=5, warm=1)
    print(json.dumps({"kernel": "laplace density head", "ms": ms, "Mpoints_s": P / ms / 1e3, "TFLOPs": 2 * 64 * 1 * 100 * P / ms / 1e9}))


def splat():
    from uncertainty_nerf_gs_b200 import binning
    from uncertainty_nerf_gs_b200.models.outputs import active_splatfacto_outputs
    G = 1_000_000
    sc = synthetic.splat_scene(G, H, W, seed=0, device=dev)
    ids, bins = binning.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], H, W)
    I = ids.numel()
    msb = timeit(lambda i: binning.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], H, W), iters=10)
    print(json.dumps({"kernel": "bin_gaussians (count+scan+expand+2 sorts+ranges)", "gaussians": G, "intersections": I,
                      "ms": msb, "Mintersections_s": I / msb / 1e3}))
    planes = [sc["rgbs"], sc["betas"], sc["depths"][:, None].contiguous()]
    ms5 = timeit(lambda i: ops.composite_tiles_planes(sc["xys"], sc["conics"], sc["opacities"], planes, ids, bins, H, W), iters=10)
    ms3 = timeit(lambda i: ops.composite_tiles(sc["xys"], sc["conics"], sc["opacities"], sc["rgbs"], ids, bins, H, W), iters=10)
    v_rgb = torch.randn(H, W, 3, device=dev); v_beta = torch.randn(H, W, 1, device=dev)
    msbw = timeit(lambda i: ops.composite_tiles_planes_backward(sc["xys"], sc["conics"], sc["opacities"],
                                                                [sc["rgbs"], sc["betas"]], ids, bins, H, W, [0.1, 0.2, 0.3, 0.0],
                                                                [v_rgb, v_beta], None), iters=5)
    ms4 = timeit(lambda i: ops.composite_tiles_planes(sc["xys"], sc["conics"], sc["opacities"], [sc["rgbs"], sc["betas"]], ids, bins, H, W), iters=10)
    print(json.dumps({"kernel": "composite_tiles backward (rgb+beta planes, 4 ch)", "ms_backward": msbw, "ms_forward_4ch": ms4}))
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    bgl = [0.1, 0.2, 0.3]
    class _BG:  # avoid the device->host sync of background.tolist() inside the timed loop
        def tolist(self): return bgl
    msf = timeit(lambda i: active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"],
                                                     sc["betas"], ids, bins, H, W, _BG()), iters=10)
    print(json.dumps({"kernel": "composite_tiles", "intersections": I, "ms_5ch_planes": ms5, "ms_3ch": ms3,
                      "ms_full_active_splatfacto_outputs": msf, "views_s": 1e3 / msf,
                      "GBs_5ch": (48 * I + 24 * R) / ms5 / 1e6, "Mpix_s": R / ms5 / 1e3}))


def cpu():
    """The CPU restatement of every kernel family on the host of this box (all threads), beside the GPU numbers
    above: bounded samples, best of 3.  The oracle is imported here only as the thing being *compared against*."""
    import time
    from oracle import compositing as oc, laplace as ol, metrics as om, reduce as orc, splat as osp
    torch.set_num_threads(os.cpu_count())
    def best(fn, n=3):
        fn()
        ts = []
        for _ in range(n):
            t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
        return min(ts)
    info = {"cpu_threads": torch.get_num_threads(), "torch": torch.__version__}
    m = synthetic.ray_samples(32768, S, seed=0)
    t = best(lambda: oc.active_nerfacto_outputs(**m))
    print(json.dumps({"cpu": "composite (oracle torch, one 32768-ray chunk)", "s": t, "Mrays_s": 32768 / t / 1e6, **info}))
    outs = synthetic.member_renders(5, H, W, seed=0, with_pred_std=True)
    t = best(lambda: orc.ensemble_reduce(outs), n=2)
    print(json.dumps({"cpu": "ensemble_reduce K=5 (oracle torch, 1297x840 image keys)", "s": t, "views_s": 1 / t, **info}))
    p_, s_, g_ = synthetic.scoring_image(800, 800, seed=0)
    t = best(lambda: om.unc_metrics_rgb(p_, g_, s_, stable=False), n=2)
    print(json.dumps({"cpu": "3x ause + auce + nll (reference's literal calls, 800x800)", "s": t, "images_s": 1 / t, **info}))
    lap = synthetic.laplace_head(32768, 64, 3, 100, seed=0)
    theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], lap["eps_draws"])
    t = best(lambda: ol.sample_laplace(lap["x"], theta, 3, torch.sigmoid))
    print(json.dumps({"cpu": "laplace rgb head moments (oracle torch loop of 100 linears, 32768 points)", "s": t,
                      "Mpoints_s": 32768 / t / 1e6, **info}))
    sc = synthetic.splat_scene(2000, 96, 128, seed=0, mean_scale_px=4.0)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], 96, 128)
    t = best(lambda: osp.rasterize(sc["xys"], sc["conics"], sc["opacities"], sc["rgbs"], ids, bins, 96, 128, torch.zeros(3)), n=1)
    print(json.dumps({"cpu": "tile rasterise 128x96, 2000 splats (oracle: python loop over tiles x splats; the reference itself "
                             "runs gsplat's CUDA kernel here, so this is not a reference timing)", "s": t,
                      "Mpix_s": 96 * 128 / t / 1e6, **info}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["composite", "backward", "reduce", "score", "laplace", "splat", "cpu"]
    for name in which:
        globals()[name]()
