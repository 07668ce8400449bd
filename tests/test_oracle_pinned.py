"""CPU: pin every oracle restatement to the REFERENCE'S OWN CODE.

Two layers (both ``-m "not gpu"``):

* ``test_live_*`` -- only where ``/root/reference`` is mounted (the dev container): the reference's methods are
  executed unmodified over the stand-in nerfstudio / gsplat packages (``oracle/ref_exec.py``) and the oracle must
  equal them BIT FOR BIT on the same inputs.
* ``test_golden_*`` -- everywhere: the oracle must equal, bit for bit, the committed ``tests/golden/ref_*.npz``
  vectors those executions produced (``tests/golden/make_golden.py``), so the pin also holds on the GPU box, where
  the reference does not exist.
"""
import os
import pathlib

import numpy as np
import pytest
import torch

from oracle import compositing as oc, laplace as ol, metrics as om, reduce as orc, ref_exec as rx, splat as osp
from uncertainty_nerf_gs_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not rx.available(), reason="/root/reference not mounted")
RAY_KEYS = ("density", "deltas", "starts", "ends", "rgb", "beta")


def same(a, b) -> bool:
    """Bit equality, NaNs in the same places counting as equal."""
    a, b = torch.as_tensor(np.asarray(a)) if not torch.is_tensor(a) else a, \
        torch.as_tensor(np.asarray(b)) if not torch.is_tensor(b) else b
    if a.shape != b.shape:
        a = a.reshape(b.shape)
    if a.is_floating_point():
        return bool(((a == b) | (a.isnan() & b.isnan())).all())
    return bool((a == b).all())


def z(name):
    return np.load(os.path.join(GOLDEN, name))


def _sys_path_stubs():
    rx.setup()
    return rx.fakes()


# ---------------------------------------------------------------------------------------------- live
@needs_ref
def test_live_weights_and_sum_modules():
    inp = synthetic.ray_samples(500, 48, seed=1)
    w = rx.compute_weights(inp["density"], inp["deltas"])               # laplace_model.py:47-62
    assert same(w, oc.get_weights(inp["density"], inp["deltas"]))
    assert same(rx.sum_module(inp["rgb"], w), torch.sum(w * inp["rgb"], dim=-2))       # :102-107
    assert same(rx.sum_module(inp["rgb"], w), oc.render_uncertainty(inp["rgb"], w))


@needs_ref
@pytest.mark.parametrize("background", ["last_sample", "white", "random"])
def test_live_active_nerfacto_get_outputs(background):
    f = _sys_path_stubs()
    inp = synthetic.ray_samples(700, 48, seed=7)
    lv = f.proposal_levels(700, 7)
    ref = rx.active_nerfacto_get_outputs(inp, lv, background=background)
    bg = (1.0, 1.0, 1.0) if background == "white" else background
    mine = oc.active_nerfacto_outputs(**inp, background_color=bg)
    assert list(ref.keys())[:9] == list(mine.keys())
    for k, v in mine.items():
        assert same(ref[k], v), k
    for i, (w, s, e) in enumerate(lv):
        assert same(ref[f"prop_depth_{i}"], oc.render_depth_median(w, s, e))


@needs_ref
def test_live_active_nerfacto_chunk_loop():
    """The inherited per-camera chunk loop: chunk-wide ``steps.min()/max()`` and the ``isnan(beta).any()`` guard
    see one chunk at a time."""
    H, W, chunk = 23, 17, 64
    inp = synthetic.ray_samples(H * W, 48, seed=3)
    ref = rx.active_nerfacto_camera(inp, H, W, chunk=chunk)
    mine = oc.render_in_chunks(oc.active_nerfacto_outputs, chunk, *[inp[k] for k in RAY_KEYS])
    for k, v in mine.items():
        assert same(ref[k], v), k


@needs_ref
def test_live_laplace_get_outputs_unc_deterministic_and_sampled():
    f = _sys_path_stubs()
    inp = synthetic.ray_samples(150, 48, seed=11, edge_cases=False)
    lv = f.proposal_levels(150, 11)
    ref = rx.laplace_get_outputs_unc(inp, None, True, lv)
    mine = oc.laplace_outputs_unc(*[inp[k] for k in RAY_KEYS])
    assert list(ref.keys())[:6] == list(mine.keys())
    for k, v in mine.items():
        assert same(ref[k], v), k
    dv = torch.rand(150, 48, 1, generator=torch.Generator().manual_seed(2)) * 0.5
    dv[::5] = 0.0
    ref = rx.laplace_get_outputs_unc(inp, dv, False, lv, seed=5)
    torch.manual_seed(5)
    std = torch.maximum(dv.sqrt(), torch.tensor([1e-10]))
    draws = torch.distributions.Normal(inp["density"], std).sample((100,))
    mine = oc.laplace_outputs_unc(*[inp[k] for k in RAY_KEYS], density_var=dv, use_deterministic_density=False,
                                  density_draws=draws)
    for k, v in mine.items():
        assert same(ref[k], v), k


@needs_ref
@pytest.mark.parametrize("out_dim,act,name", [(3, torch.nn.Sigmoid(), "sigmoid"), (1, torch.exp, "exp")])
def test_live_sample_laplace(out_dim, act, name):
    lap = synthetic.laplace_head(300, 64, out_dim, 100, seed=out_dim)
    lin = torch.nn.Linear(64, out_dim)
    with torch.no_grad():
        lin.weight.copy_(lap["mu_q"][:64 * out_dim].view(out_dim, 64))
        lin.bias.copy_(lap["mu_q"][64 * out_dim:])
    mu, s2 = rx.sample_laplace(lin, act, lap["ggn"], lap["x"], 100, 2.0, 1e-9, seed=4)   # laplace_field.py:528-568
    torch.manual_seed(4)
    theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], torch.randn(100, 64 * out_dim + out_dim), prior_prec=2.0)
    m, _, ss = ol.sample_laplace(lap["x"], theta, out_dim, torch.sigmoid if name == "sigmoid" else torch.exp)
    assert same(mu, m) and same(s2, ss)
    assert same(lin.weight.flatten(), lap["mu_q"][:64 * out_dim])      # MAP parameters restored (:567)


@needs_ref
@pytest.mark.parametrize("k,seed,pred_std", [(5, 3, False), (3, 4, True), (2, 6, True)])
def test_live_ensemble_reduce(k, seed, pred_std):
    outs = synthetic.member_renders(k, 9, 11, seed=seed, with_pred_std=pred_std)
    ref, mine = rx.ensemble_reduce(outs), orc.ensemble_reduce(outs)       # ensemble_pipeline.py:159-190
    assert list(ref.keys()) == list(mine.keys())
    assert all(same(ref[kk], mine[kk]) for kk in ref)


@needs_ref
def test_live_mcdropout_reduce():
    outs = synthetic.member_renders(10, 9, 11, seed=5)
    ref, mine = rx.mcdropout_reduce(outs), orc.mcdropout_reduce(outs)     # mcdropout_models.py:114-126
    assert list(ref.keys()) == list(mine.keys())
    assert all(same(ref[kk], mine[kk]) for kk in ref)


@needs_ref
def test_live_rgb_scoring_and_nll():
    p, s, g = synthetic.scoring_image(60, 70, seed=3)
    ref = rx.unc_metrics_rgb({"rgb": p, "rgb_std": s}, g)                 # eval_uncertainty.py:306-402
    mine = om.unc_metrics_rgb(p, g, s)
    for k, v in mine.items():
        assert same(np.asarray(ref[k], dtype=np.float64) if not torch.is_tensor(ref[k]) else ref[k],
                    np.asarray(v, dtype=np.float64) if not torch.is_tensor(v) else v), k
    pro = om.rgb_metric_prologue(p, g, s)
    assert same(ref["mse"], pro["squared_error"]) and same(ref["absolute_error"], pro["absolute_error"])
    assert same(rx.nll(p.reshape(-1, 3), g.reshape(-1, 3), s, 3e-2),
                om.negative_gaussian_loglikelihood(p.reshape(-1, 3), g.reshape(-1, 3), s, 3e-2))   # :404-412


def _real_torchvision() -> bool:
    try:
        import torchvision.transforms.functional as F
        return hasattr(F, "resize") and callable(F.resize) and "torchvision" in (getattr(F.resize, "__module__", "") or "")
    except Exception:
        return False


@needs_ref
@pytest.mark.parametrize("render_hw", [(40, 50), (39, 49)])     # (39, 49): splatfacto's [H-1, W-1] depth -> the resize branch
def test_live_depth_scoring(tmp_path, render_hw):
    if render_hw != (40, 50) and not _real_torchvision():
        pytest.skip("the resize branch calls torchvision's F.resize: needs the real package")
    gen = torch.Generator().manual_seed(1)
    d = torch.rand(*render_hw, 1, generator=gen) * 4
    ds = torch.rand(*render_hw, 1, generator=gen) * 0.3 + 0.01
    gt = torch.rand(40, 50, generator=gen) * 5
    gt[gt < 0.7] = 0
    rx.write_depth_side_inputs(tmp_path / "data", [gt.numpy()], 1.7)
    ref = rx.unc_metrics_depth({"depth": d, "depth_std": ds}, 0, tmp_path / "data", tmp_path)     # :415-644
    scale = float(np.loadtxt(str(tmp_path / "data") + "/scale_parameters.txt", delimiter=","))
    mine = om.unc_metrics_depth(d, ds, gt, scale)
    for k, v in mine.items():
        assert same(np.asarray(ref[k], dtype=np.float64) if not torch.is_tensor(ref[k]) else ref[k],
                    np.asarray(v, dtype=np.float64) if not torch.is_tensor(v) else v), k


@needs_ref
def test_live_test_set_loop_and_npy_dumps(tmp_path):
    """``get_average_uncertainty_metrics`` end to end: per-image scalars -> float32 means; curves -> the
    ``auce_rgb_*.npy`` files ``plot_auce_curves`` writes (metrics/auce.py:130-141)."""
    views, gts = [], []
    for i in range(3):
        p, s, g = synthetic.scoring_image(30, 40, seed=20 + i)
        views.append({"rgb": p, "rgb_std": s, "accumulation": torch.ones(30, 40, 1), "depth": torch.ones(30, 40, 1)})
        gts.append(g)
    res = rx.average_uncertainty_metrics(views, gts, tmp_path, image_metrics=(20.0, 0.9, 0.1))
    per = []
    curves = []
    for v, g in zip(views, gts):
        d = om.unc_metrics_rgb(v["rgb"], g, v["rgb_std"])
        per.append(om.per_image_rgb_scalars(d))
        curves.append(d)
    agg = om.aggregate_scalars(per)                                        # :1070-1077
    for k, v in agg.items():
        assert res[k] == v, k
    assert list(res.keys())[:3] == ["psnr", "ssim", "lpips"] and list(res.keys())[-2:] == ["num_rays_per_sec", "fps"]
    names = {"coverage_values": "empirical_coverage", "avg_length_values": "avg_length",
             "coverage_error_values": "empirical_coverage_error",
             "abs_coverage_error_values": "empirical_coverage_absolute_error",
             "neg_coverage_error_values": "empirical_coverage_negative_error"}
    for key, fn in names.items():
        want = om.aggregate_curves([c[key] for c in curves])               # :920-946, 1040-1046
        got = np.load(tmp_path / "plots" / f"auce_rgb_{fn}.npy")
        assert np.array_equal(got, want), key
    assert np.array_equal(np.load(tmp_path / "plots" / "auce_rgb_alphas.npy"), np.array(om.auce_alphas()))


@needs_ref
def test_live_active_splatfacto_get_outputs():
    """``ActiveSplatfactoModel.get_outputs(camera)`` (activesplatfacto_model.py:142-367) against the oracle's
    pass structure.  gsplat itself is the stand-in (= oracle.splat), so this pins everything AROUND the
    rasteriser: camera conventions, SH colours, activations, the four passes, normalisation, residual gather."""
    rx.setup()
    from nerfstudio.cameras.cameras import Cameras

    H, W, G = 40, 56, 300
    sc, cam, gauss, log_unc, c2w = _splat_case(H, W, G)
    ref = rx.active_splatfacto_get_outputs(gauss, log_unc, Cameras(c2w[:3], sc["fx"], sc["fy"], sc["cx"], sc["cy"], W, H))
    mine = _oracle_splat(sc, c2w, gauss, log_unc, H, W)
    for k, v in mine.items():
        assert same(ref[k], v), k
    assert list(ref.keys())[:9] == list(mine.keys())


def _splat_case(H, W, G, seed=2):
    sc = synthetic.gaussians_3d(G, H, W, seed=seed, sh_degree=3)
    vm = torch.eye(4)
    vm[:3] = sc["viewmat"]
    c2w = torch.linalg.inv(vm)
    c2w[:3, :3] = c2w[:3, :3] @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))
    gauss = {"means": sc["means"], "scales": torch.log(sc["scales"]), "quats": sc["quats"],
             "features_dc": sc["sh_coeffs"][:, 0, :], "features_rest": sc["sh_coeffs"][:, 1:, :],
             "opacities": torch.logit(sc["opacities"])}
    log_unc = torch.randn(G, 1, generator=torch.Generator().manual_seed(9))
    return sc, None, gauss, log_unc, c2w


def _oracle_splat(sc, c2w, gauss, log_unc, H, W, background=(0.1, 0.2, 0.3)):
    R = c2w[:3, :3] @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))
    T = c2w[:3, 3:4]
    viewmat = torch.eye(4)
    viewmat[:3, :3] = R.T
    viewmat[:3, 3:4] = -R.T @ T
    fx, fy, cx, cy = (torch.tensor(float(sc[k])).float().item() for k in ("fx", "fy", "cx", "cy"))
    p = osp.project_gaussians(gauss["means"], torch.exp(gauss["scales"]), 1,
                              gauss["quats"] / gauss["quats"].norm(dim=-1, keepdim=True), viewmat[:3, :], fx, fy, cx, cy,
                              H, W)
    colors = torch.cat((gauss["features_dc"][:, None, :], gauss["features_rest"]), dim=1)
    rgbs = torch.clamp(osp.spherical_harmonics(3, gauss["means"] - c2w[:3, 3], colors) + 0.5, min=0.0)
    betas = torch.nn.functional.softplus(log_unc) + 0.01
    ids, bins = osp.bin_gaussians(p["xys"], p["depths"], p["radii"], H, W)
    return osp.active_splatfacto_outputs(p["xys"], p["depths"], p["conics"], torch.sigmoid(gauss["opacities"]), rgbs,
                                         betas, ids, bins, H, W, torch.tensor(background))


# ---------------------------------------------------------------------------------------------- golden
def test_golden_composite():
    g = z("ref_composite.npz")
    R, S, seed, H, W, chunk, cseed = (int(v) for v in g["meta"])
    inp = synthetic.ray_samples(R, S, seed=seed)
    for tag, bg in (("eval", "last_sample"), ("white", (1.0, 1.0, 1.0))):
        mine = oc.active_nerfacto_outputs(**inp, background_color=bg)
        for k, v in mine.items():
            if k != "density":
                assert same(g[f"{tag}_{k}"], v), (tag, k)
    w = oc.get_weights(inp["density"], inp["deltas"])
    assert same(g["train_weights"], w)
    assert same(g["train_rgb"], oc.render_rgb(inp["rgb"], w, "last_sample", training=True))
    inp_c = synthetic.ray_samples(H * W, S, seed=cseed)
    mine = oc.render_in_chunks(oc.active_nerfacto_outputs, chunk, *[inp_c[k] for k in RAY_KEYS])
    for k, v in mine.items():
        if k != "density":
            assert same(g[f"camera_{k}"], v), k


def test_golden_laplace():
    g = z("ref_laplace.npz")
    inp = synthetic.ray_samples(80, 48, seed=11, edge_cases=False)
    for k, v in oc.laplace_outputs_unc(*[inp[k] for k in RAY_KEYS]).items():
        assert same(g[f"det_{k}"], v), k
    inp_s = synthetic.ray_samples(16, 48, seed=13, edge_cases=False)
    mine = oc.laplace_outputs_unc(*[inp_s[k] for k in RAY_KEYS], density_var=torch.from_numpy(g["density_var"]),
                                  use_deterministic_density=False, density_draws=torch.from_numpy(g["density_draws"]))
    for k, v in mine.items():
        assert same(g[f"samp_{k}"], v), k
    for head, od, act in (("rgb", 3, torch.sigmoid), ("density", 1, torch.exp)):
        lap = synthetic.laplace_head(257, 64, od, 100, seed=5 + od)
        theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], torch.from_numpy(g[f"{head}_randn"]))
        m, _, s2 = ol.sample_laplace(lap["x"], theta, od, act)
        assert same(g[f"{head}_mu"], m) and same(g[f"{head}_sigma2"], s2), head


def test_golden_reduce():
    g = z("ref_reduce.npz")
    for tag, (k, seed, std, fn) in {"ensB": (5, 3, False, orc.ensemble_reduce), "ensA": (3, 4, True, orc.ensemble_reduce),
                                    "mcd": (10, 5, False, orc.mcdropout_reduce)}.items():
        r = fn(synthetic.member_renders(k, 9, 11, seed=seed, with_pred_std=std))
        assert list(r.keys()) == list(g[f"{tag}_keys"])
        for kk, v in r.items():
            assert same(g[f"{tag}_{kk}"], v), (tag, kk)


def test_golden_scoring():
    g = z("ref_scoring.npz")
    p, s, gt = synthetic.scoring_image(60, 70, seed=3)
    mine = om.unc_metrics_rgb(p, gt, s)
    for k, v in mine.items():
        assert same(g[f"rgb_{k}"], np.asarray(v, dtype=np.float64) if not torch.is_tensor(v) else v.double()), k
    d = om.unc_metrics_depth(torch.from_numpy(g["depth_in"]), torch.from_numpy(g["depth_std_in"]),
                             torch.from_numpy(g["depth_gt_in"]), float(g["depth_scale"]))
    for k, v in d.items():
        assert same(g[f"depth_{k}"], np.asarray(v, dtype=np.float64) if not torch.is_tensor(v) else v), k


def test_golden_splat_outputs_cover_the_reference_keys():
    g = z("ref_splat.npz")
    H, W, G, seed, deg = (int(v) for v in g["meta"])
    sc, _, gauss, log_unc, c2w = _splat_case(H, W, G, seed)
    assert same(c2w[:3], g["c2w"]) and same(log_unc, g["log_unc"])
    mine = _oracle_splat(sc, c2w, gauss, log_unc, H, W)
    for k, v in mine.items():
        assert same(g[f"out_{k}"], v), k
