"""Generate the committed golden vectors.  Run in the dev container, where ``/root/reference`` is mounted:

    python tests/golden/make_golden.py

* ``ause_golden.npz`` / ``auce_golden.npz`` -- outputs of the REFERENCE's own
  ``nerfuncertainty/metrics/ause.py`` and ``auce.py`` (imported by file path, ``oracle/ref_loader.py``)
  on small seeded inputs.  ``ause`` runs with ``torch.sort`` forced to ``stable=True`` (the ranking
  contract); ``ause_nostable`` entries record the literal default call on tie-free inputs, where the two
  coincide.
* ``composite_golden.npz`` / ``reduce_golden.npz`` / ``laplace_golden.npz`` / ``splat_golden.npz`` -- outputs of the
  oracle restatements, kept to detect drift of the oracle across torch versions.
* ``ref_*.npz`` -- outputs of the REFERENCE'S OWN model / scoring methods, executed unmodified on CPU over the
  stand-in nerfstudio / gsplat packages (``oracle/ref_exec.py``): ``ActiveNerfactoModel.get_outputs`` (eval,
  training, chunk loop), ``NerfactoLaplaceModel.get_outputs_unc`` (deterministic and sampled density),
  ``NerfactoLaplaceField.sample_laplace``, the MC-dropout and ensemble reduces, ``ActiveSplatfactoModel.get_outputs``,
  ``get_unc_metrics_rgb`` / ``get_unc_metrics_depth`` and the ``get_average_uncertainty_metrics`` loop with its
  ``auce_*.npy`` dumps.  The CUDA path is compared with these on the GPU box (``tests/test_gpu_reference_golden.py``).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import compositing as oc, laplace as ol, reduce as orc, ref_loader, splat as osp  # noqa: E402
from uncertainty_nerf_gs_b200 import synthetic  # noqa: E402


def _np(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items() if v is not None}


def make_reference_executed():
    """``ref_*.npz``: see the module docstring.  Inputs that ``synthetic`` regenerates from a seed are not stored."""
    import pathlib
    import tempfile

    from oracle import ref_exec as rx

    rx.setup()
    f = rx.fakes()
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **kw)

    # ---- active-nerfacto: one chunk (eval + training), fixed background, and the chunk loop ----
    R = 96
    inp = synthetic.ray_samples(R, 48, seed=7)
    lv = f.proposal_levels(R, 7)
    out = {}
    for tag, kw in (("eval", {}), ("white", {"background": "white"}), ("train", {"training": True})):
        o = rx.active_nerfacto_get_outputs(inp, lv, **kw)
        if tag == "train":
            o["weights"] = o["weights_list"][-1]
        out.update({f"{tag}_{k}": v for k, v in _np({k: v for k, v in o.items() if torch.is_tensor(v) and k != "density"}).items()})
    H, W, chunk = 12, 10, 32
    inp_c = synthetic.ray_samples(H * W, 48, seed=3)
    oc_ = rx.active_nerfacto_camera(inp_c, H, W, chunk=chunk)
    out.update({f"camera_{k}": v for k, v in _np(oc_).items() if k != "density"})
    save("ref_composite.npz", **out, meta=np.array([R, 48, 7, H, W, chunk, 3]))

    # ---- nerfacto-laplace: compositing with deterministic and sampled density; sample_laplace for both heads ----
    inp = synthetic.ray_samples(80, 48, seed=11, edge_cases=False)
    lv = f.proposal_levels(80, 11)
    out = {f"det_{k}": v for k, v in _np(rx.laplace_get_outputs_unc(inp, None, True, lv)).items()}
    inp_s = synthetic.ray_samples(16, 48, seed=13, edge_cases=False)      # small: the 100 draws are stored
    lv_s = f.proposal_levels(16, 13)
    g = torch.Generator().manual_seed(12)
    dv = torch.rand(16, 48, 1, generator=g) * 0.5
    dv[::7] = 0.0                                             # std clamped to 1e-10 (laplace_model.py:491)
    samp = rx.laplace_get_outputs_unc(inp_s, dv, False, lv_s, seed=5)
    torch.manual_seed(5)                                      # the draws the reference just consumed
    std = torch.maximum(dv.sqrt(), torch.tensor([1e-10]))
    draws = torch.distributions.Normal(inp_s["density"], std).sample((100,))
    out.update({f"samp_{k}": v for k, v in _np(samp).items()})
    out["density_var"] = dv.numpy()
    out["density_draws"] = draws.numpy().astype(np.float32)   # [100, 16, 48, 1]
    for head, (od, act, name) in {"rgb": (3, torch.nn.Sigmoid(), "sigmoid"), "density": (1, torch.exp, "exp")}.items():
        lap = synthetic.laplace_head(257, 64, od, 100, seed=5 + od)
        lin = torch.nn.Linear(64, od)
        with torch.no_grad():
            lin.weight.copy_(lap["mu_q"][:64 * od].view(od, 64))
            lin.bias.copy_(lap["mu_q"][64 * od:])
        mu, s2 = rx.sample_laplace(lin, act, lap["ggn"], lap["x"], 100, 1.0, 1e-9, seed=od)
        torch.manual_seed(od)
        out[f"{head}_randn"] = torch.randn(100, 64 * od + od).numpy()
        out[f"{head}_mu"], out[f"{head}_sigma2"] = mu.numpy(), s2.numpy()
    save("ref_laplace.npz", **out)

    # ---- reduces (inputs: synthetic.member_renders, regenerated from the seeds below) ----
    out = {}
    for tag, (k, seed, std, fn) in {"ensB": (5, 3, False, rx.ensemble_reduce), "ensA": (3, 4, True, rx.ensemble_reduce),
                                    "mcd": (10, 5, False, rx.mcdropout_reduce)}.items():
        r = fn(synthetic.member_renders(k, 9, 11, seed=seed, with_pred_std=std))
        out[f"{tag}_keys"] = np.array(list(r.keys()))
        out.update({f"{tag}_{kk}": v for kk, v in _np(r).items()})
    save("ref_reduce.npz", **out)

    # ---- scoring: rgb, depth, and the test-set loop with its .npy dumps ----
    out = {}
    p, s, g_ = synthetic.scoring_image(60, 70, seed=3)
    r = rx.unc_metrics_rgb({"rgb": p, "rgb_std": s}, g_)
    out.update({f"rgb_{k}": np.asarray(v, dtype=np.float64) for k, v in r.items()
                if not torch.is_tensor(v) or k == "mse"})
    out["rgb_nll_map"] = r["neg_log_prob"].numpy()
    with tempfile.TemporaryDirectory() as td:
        td = pathlib.Path(td)
        gen = torch.Generator().manual_seed(1)
        d = torch.rand(40, 50, 1, generator=gen) * 4
        ds = torch.rand(40, 50, 1, generator=gen) * 0.3 + 0.01
        gt = torch.rand(40, 50, generator=gen) * 5
        gt[gt < 0.7] = 0
        rx.write_depth_side_inputs(td / "data", [gt.numpy()], 1.7)
        scale = float(np.loadtxt(str(td / "data") + "/scale_parameters.txt", delimiter=","))
        r = rx.unc_metrics_depth({"depth": d, "depth_std": ds}, 0, td / "data", td)
        out.update(depth_in=d.numpy(), depth_std_in=ds.numpy(), depth_gt_in=gt.numpy(), depth_scale=np.float64(scale))
        out.update({f"depth_{k}": np.asarray(v, dtype=np.float64) for k, v in r.items()
                    if not torch.is_tensor(v)})
        out["depth_mse"] = r["mse"].numpy()
        # test-set loop over 3 views with rgb + depth scoring
        views, gts, dgts = [], [], []
        for i in range(3):
            pr, st, gi = synthetic.scoring_image(30, 40, seed=20 + i)
            gen = torch.Generator().manual_seed(30 + i)
            di = torch.rand(30, 40, 1, generator=gen) * 4
            dsi = torch.rand(30, 40, 1, generator=gen) * 0.3 + 0.01
            dg = torch.rand(30, 40, generator=gen) * 5
            dg[dg < 0.5] = 0
            views.append({"rgb": pr, "rgb_std": st, "accumulation": torch.ones(30, 40, 1), "depth": di, "depth_std": dsi})
            gts.append(gi)
            dgts.append(dg.numpy())
        rx.write_depth_side_inputs(td / "set", dgts, 0.9)
        res = rx.average_uncertainty_metrics(views, gts, td / "run", dataset_path=td / "set", eval_depth=True,
                                             image_metrics=(20.0, 0.9, 0.1), min_depth_std_for_nll=2.0)
        out["set_keys"] = np.array(list(res.keys()))
        out["set_values"] = np.array([res[k] for k in res.keys()], dtype=np.float64)
        out["set_depth_gt"] = np.stack(dgts)
        for fn in sorted(os.listdir(td / "run" / "plots")):
            if fn.endswith(".npy"):
                out["npy_" + fn[:-4]] = np.load(td / "run" / "plots" / fn)
    save("ref_scoring.npz", **out)

    # ---- active-splatfacto: the reference's get_outputs(camera) on a 3-D scene ----
    from nerfstudio.cameras.cameras import Cameras

    H, W, G = 40, 56, 500
    sc = synthetic.gaussians_3d(G, H, W, seed=2, sh_degree=3)
    vm = torch.eye(4)
    vm[:3] = sc["viewmat"]
    c2w = torch.linalg.inv(vm)
    c2w[:3, :3] = c2w[:3, :3] @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))     # gsplat -> nerfstudio camera axes
    cam = Cameras(c2w[:3], sc["fx"], sc["fy"], sc["cx"], sc["cy"], W, H)
    gauss = {"means": sc["means"], "scales": torch.log(sc["scales"]), "quats": sc["quats"],
             "features_dc": sc["sh_coeffs"][:, 0, :], "features_rest": sc["sh_coeffs"][:, 1:, :],
             "opacities": torch.logit(sc["opacities"])}
    log_unc = torch.randn(G, 1, generator=torch.Generator().manual_seed(9))
    o = rx.active_splatfacto_get_outputs(gauss, log_unc, cam)
    save("ref_splat.npz", c2w=c2w[:3].numpy(), log_unc=log_unc.numpy(), meta=np.array([H, W, G, 2, 3]),
         **{f"out_{k}": v for k, v in _np(o).items()})


def main():
    loaded = ref_loader.load_reference_metrics()
    if loaded is None:
        raise SystemExit("/root/reference is not mounted: cannot regenerate reference-derived goldens")
    ref_ause, ref_auce = loaded
    torch.manual_seed(0)

    # ---- AUSE: ties at a variance floor, N = 10007 ----
    g = torch.Generator().manual_seed(1234)
    n = 10007
    unc = (torch.clamp(0.1 * torch.rand(n, generator=g), min=0.03) ** 2)
    err = torch.rand(n, generator=g) ** 2
    out = {"unc": unc.numpy(), "err": err.numpy()}
    with ref_loader.stable_torch_sort():
        for et in ("mae", "mse", "rmse"):
            ratio, e, v, a = ref_ause(unc, err, et)
            out[f"{et}_err"] = np.asarray(e, dtype=np.float64)
            out[f"{et}_err_by_var"] = np.asarray(v, dtype=np.float64)
            out[f"{et}_ause"] = np.float64(a)
    out["ratio"] = ratio
    # tie-free input: the reference's literal (unstable) call is reproducible
    unc_nt = (torch.randperm(n, generator=g).float() + 1.0) / n
    assert unc_nt.unique().numel() == n
    out["unc_notie"] = unc_nt.numpy()
    for et in ("mae", "rmse"):
        _, e, v, a = ref_ause(unc_nt, err, et)
        out[f"notie_{et}_err"] = np.asarray(e, dtype=np.float64)
        out[f"notie_{et}_err_by_var"] = np.asarray(v, dtype=np.float64)
        out[f"notie_{et}_ause"] = np.float64(a)
    np.savez_compressed(os.path.join(HERE, "ause_golden.npz"), **out)

    # ---- AUCE: 4096 x 3 float32 with exact hits, sigma = 0, NaN targets ----
    g = torch.Generator().manual_seed(99)
    m = torch.rand(4096, 3, generator=g)
    s = torch.clamp(0.1 * torch.rand(4096, 1, generator=g), min=0.03).repeat(1, 3)
    t = torch.clamp(m + s * torch.randn(4096, 3, generator=g), 0, 1)
    t[::17] = m[::17]
    s[5::29] = 0.0
    t[7::31] = float("nan")
    d = ref_auce(m.numpy(), s.numpy(), t.numpy())
    np.savez_compressed(os.path.join(HERE, "auce_golden.npz"), mean=m.numpy(), sigma=s.numpy(), target=t.numpy(),
                        **{k: np.asarray(v) for k, v in d.items()})

    # ---- oracle-derived goldens ----
    inp = synthetic.ray_samples(64, 48, seed=7)
    ref = oc.active_nerfacto_outputs(**inp)
    np.savez_compressed(os.path.join(HERE, "composite_golden.npz"),
                        **{f"in_{k}": v.numpy() for k, v in inp.items()},
                        **{f"out_{k}": v.numpy() for k, v in ref.items() if k != "density"},
                        out_weights=oc.get_weights(inp["density"], inp["deltas"]).numpy())

    outs = synthetic.member_renders(5, 9, 11, seed=3)
    red = orc.ensemble_reduce(outs)
    np.savez_compressed(os.path.join(HERE, "reduce_golden.npz"), **{f"out_{k}": v.numpy() for k, v in red.items()})

    lap = synthetic.laplace_head(257, 64, 3, 100, seed=5)
    theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], lap["eps_draws"])
    mu, mu2, s2 = ol.sample_laplace(lap["x"], theta, 3, torch.sigmoid)
    np.savez_compressed(os.path.join(HERE, "laplace_golden.npz"), mean=mu.numpy(), mean2=mu2.numpy(),
                        sigma2=s2.numpy())

    sc = synthetic.splat_scene(400, 40, 56, seed=2, mean_scale_px=4.0)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], 40, 56)
    so = osp.active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"],
                                       sc["betas"], ids, bins, 40, 56, torch.tensor([0.1, 0.2, 0.3]))
    np.savez_compressed(os.path.join(HERE, "splat_golden.npz"), ids=ids.numpy(), bins=bins.numpy(),
                        **{f"out_{k}": v.numpy() for k, v in so.items()})
    make_reference_executed()
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
