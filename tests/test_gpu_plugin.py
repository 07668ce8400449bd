"""GPU: the plugin replacements (``models/nerfstudio_plugin.py``) EXECUTE, bound to stand-in model objects built on
the stand-in nerfstudio package (``tests/stubs``), and reproduce what the reference's own methods returned for the
same inputs (``tests/golden/ref_*.npz``: the reference executed unmodified in the dev container,
``tests/golden/make_golden.py``).  Neither nerfstudio nor the reference exists on the GPU box.
"""
import os
import sys
import types
from dataclasses import dataclass

import numpy as np
import pytest
import torch

from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
RAY_KEYS = ("density", "deltas", "starts", "ends", "rgb", "beta")
ATOL = {"rgb": 2e-6, "accumulation": 2e-6, "expected_depth": 2e-6, "rgb_var": 1e-7, "rgb_std": 1e-6,
        "depth_var": 1e-7, "depth_std": 1e-6, "weights": 2e-7}


@pytest.fixture(scope="module")
def stubs(built_library):
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    import ub_stubs

    ub_stubs.install()
    import fakes

    return fakes


def z(name):
    return np.load(os.path.join(GOLDEN, name))


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _close(got, want, key, rtol=1e-5):
    want = torch.from_numpy(np.asarray(want))
    torch.testing.assert_close(got.cpu().reshape(want.shape), want, rtol=rtol, atol=ATOL.get(key, 1e-6), equal_nan=True,
                               msg=lambda m: f"{key}: {m}")


def _depth_close(got, want, max_flips=1):
    """Median depth is an index decision: equal except where the cumulative weight sits within an ulp of 0.5."""
    want = torch.from_numpy(np.asarray(want))
    same = torch.isclose(got.cpu().reshape(want.shape), want, rtol=1e-6, atol=0.0)
    assert int((~same).sum()) <= max_flips
    return same.reshape(-1)


def _active_nerfacto_standin(fakes, inp, levels, background="last_sample", training=False, chunk=1 << 15):
    from nerfstudio.models.nerfacto import NerfactoModel, NerfactoModelConfig

    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    class StandIn(NerfactoModel):
        get_outputs = plug.active_nerfacto_get_outputs

    cfg = NerfactoModelConfig(background_color=background, eval_num_rays_per_chunk=chunk,
                              num_proposal_iterations=len(levels))
    model = StandIn(cfg).cuda()
    fakes.attach_producers(model, inp, levels)
    model.train(training)
    return model


def test_active_nerfacto_get_outputs_eval_and_backgrounds(stubs):
    g = z("ref_composite.npz")
    R, S, seed = (int(v) for v in g["meta"][:3])
    inp = _cuda(synthetic.ray_samples(R, S, seed=seed))
    lv = [tuple(t.cuda() for t in l) for l in stubs.proposal_levels(R, seed)]
    for tag, bg in (("eval", "last_sample"), ("white", "white")):
        model = _active_nerfacto_standin(stubs, inp, lv, background=bg)
        with torch.no_grad():
            out = model.get_outputs(stubs.flat_ray_bundle(R, device="cuda"))
        assert list(out.keys()) == ["rgb", "accumulation", "depth", "expected_depth", "density", "rgb_var", "rgb_std",
                                    "depth_var", "depth_std", "prop_depth_0", "prop_depth_1"]
        ok = _depth_close(out["depth"], g[f"{tag}_depth"])
        for k in ("rgb", "accumulation", "expected_depth", "rgb_var", "rgb_std"):
            _close(out[k], g[f"{tag}_{k}"], k)
        for k in ("depth_var", "depth_std"):
            _close(out[k][ok], g[f"{tag}_{k}"][ok.numpy()], k)
        for i in range(2):
            assert torch.equal(out[f"prop_depth_{i}"].cpu(), torch.from_numpy(g[f"{tag}_prop_depth_{i}"]))
        assert torch.equal(out["density"], inp["density"])                   # passed through, as in the reference


def test_active_nerfacto_get_outputs_with_derived_deltas(stubs, monkeypatch):
    """``UB_DERIVE_DELTAS=1``: the patched ``get_outputs`` hands the compositor no deltas; for ray samples built the way
    ``RayBundle.get_ray_samples`` builds them (deltas = ends - starts) every output is bit for bit the default path's."""
    R, S = 777, 48
    inp = _cuda(synthetic.ray_samples(R, S, seed=5))
    inp["deltas"] = inp["ends"] - inp["starts"]
    lv = [tuple(t.cuda() for t in l) for l in stubs.proposal_levels(R, 5)]
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("UB_DERIVE_DELTAS", flag)
        model = _active_nerfacto_standin(stubs, inp, lv, chunk=256)
        with torch.no_grad():
            outs.append(model.get_outputs(stubs.flat_ray_bundle(R, device="cuda")))
    assert list(outs[0].keys()) == list(outs[1].keys())
    for k in outs[0]:
        assert torch.equal(outs[0][k].view(torch.int32), outs[1][k].view(torch.int32)), k


def test_active_nerfacto_camera_chunk_loop(stubs):
    """The inherited ``get_outputs_for_camera_ray_bundle`` calls the replacement once per eval chunk."""
    g = z("ref_composite.npz")
    S, H, W, chunk, cseed = int(g["meta"][1]), *(int(v) for v in g["meta"][3:7])
    inp = _cuda(synthetic.ray_samples(H * W, S, seed=cseed))
    model = _active_nerfacto_standin(stubs, inp, [], chunk=chunk)
    out = model.get_outputs_for_camera_ray_bundle(stubs.camera_ray_bundle(H, W, device="cuda"))
    ok = _depth_close(out["depth"], g["camera_depth"])
    for k in ("rgb", "accumulation", "expected_depth", "rgb_var", "rgb_std"):
        assert out[k].shape[:2] == (H, W)
        _close(out[k], g[f"camera_{k}"], k)
    _close(out["depth_var"].reshape(-1, 1)[ok], g["camera_depth_var"].reshape(-1, 1)[ok.numpy()], "depth_var")


def test_active_nerfacto_training_mode_uses_the_fused_backward(stubs):
    """``self.training``: the same kernel through ``autograd.composite_rays_train`` -- unclamped colours, weights and
    sample lists in the output dict like the reference, gradients to the field outputs."""
    g = z("ref_composite.npz")
    R, S, seed = (int(v) for v in g["meta"][:3])
    inp = _cuda(synthetic.ray_samples(R, S, seed=seed))
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("density", "rgb", "beta")}
    inp_t = dict(inp, **leaves)
    lv = [tuple(t.cuda() for t in l) for l in stubs.proposal_levels(R, seed)]
    model = _active_nerfacto_standin(stubs, inp_t, lv, training=True)
    out = model.get_outputs(stubs.flat_ray_bundle(R, device="cuda"))
    assert list(out.keys()) == ["rgb", "accumulation", "depth", "expected_depth", "density", "rgb_var", "rgb_std",
                                "depth_var", "depth_std", "weights_list", "ray_samples_list", "prop_depth_0", "prop_depth_1"]
    assert len(out["weights_list"]) == 3 and len(out["ray_samples_list"]) == 3
    _close(out["weights_list"][-1], g["train_weights"], "weights")
    for k in ("rgb", "accumulation", "expected_depth", "rgb_var"):
        _close(out[k], g[f"train_{k}"], k)
    loss = out["rgb"].sum() + out["rgb_var"].sum() + out["accumulation"].sum() + out["weights_list"][-1].pow(2).sum()
    loss.backward()
    # the same loss through torch autograd on the CPU oracle, with the reference's guard (`nan_to_num` if any NaN)
    from oracle import compositing as oc

    cpu = synthetic.ray_samples(R, S, seed=seed)
    ref_leaves = {k: cpu[k].clone().requires_grad_(True) for k in ("density", "rgb", "beta")}
    beta = torch.nan_to_num(ref_leaves["beta"], 0.0)
    w = oc.get_weights(ref_leaves["density"], cpu["deltas"])
    ref_loss = (oc.render_rgb(ref_leaves["rgb"], w, "last_sample", training=True).sum() + oc.render_uncertainty(beta, w ** 2).sum()
                + oc.render_accumulation(w).sum() + w.pow(2).sum())
    ref_loss.backward()
    for k, v in leaves.items():
        want = ref_leaves[k].grad
        scale = max(1.0, float(torch.nan_to_num(want).abs().max()))
        torch.testing.assert_close(v.grad.cpu(), want, rtol=2e-4, atol=2e-6 * scale, equal_nan=True, msg=lambda m: f"grad {k}: {m}")
    # the NaN betas of the synthetic input were replaced (stability guard) and receive no gradient
    nan_mask = torch.isnan(inp["beta"])
    assert bool(nan_mask.any()) and float(leaves["beta"].grad[nan_mask].abs().sum()) == 0.0


def _laplace_standin(fakes, inp, density_var, levels):
    from nerfstudio.models.nerfacto import NerfactoModel, NerfactoModelConfig

    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    class StandIn(NerfactoModel):
        get_outputs_unc = plug.laplace_get_outputs_unc

    model = StandIn(NerfactoModelConfig(num_proposal_iterations=len(levels))).cuda()
    field = fakes.TensorField(inp["density"], inp["rgb"], rgb_var=inp["beta"], density_var=density_var)
    fakes.attach_producers(model, inp, levels, field=field)
    model.eval()
    return model


def test_laplace_get_outputs_unc(stubs, monkeypatch):
    g = z("ref_laplace.npz")
    inp = _cuda(synthetic.ray_samples(80, 48, seed=11, edge_cases=False))
    lv = [tuple(t.cuda() for t in l) for l in stubs.proposal_levels(80, 11)]
    model = _laplace_standin(stubs, inp, None, lv)
    out = model.get_outputs_unc(stubs.flat_ray_bundle(80, device="cuda"), is_inference=True, use_deterministic_density=True,
                                prior_prec=3.0, n_samples=7)
    assert list(out.keys()) == ["rgb", "rgb_std", "accumulation", "depth", "depth_std", "expected_depth", "prop_depth_0",
                                "prop_depth_1"]
    assert model.field.calls[-1]["prior_prec"] == 3.0 and model.field.calls[-1]["n_samples"] == 7   # forwarded like :459-465
    ok = _depth_close(out["depth"], g["det_depth"])
    for k in ("rgb", "rgb_std", "accumulation", "expected_depth"):
        _close(out[k], g[f"det_{k}"], k)
    _close(out["depth_std"][ok], g["det_depth_std"][ok.numpy()], "depth_std")
    # sampled density, draw-for-draw: hand the replacement the reference's own 100 draws through torch.randn
    inp_s = _cuda(synthetic.ray_samples(16, 48, seed=13, edge_cases=False))
    lv_s = [tuple(t.cuda() for t in l) for l in stubs.proposal_levels(16, 13)]
    dv = torch.from_numpy(g["density_var"]).cuda()
    std = torch.maximum(dv.sqrt(), torch.tensor(1e-10, device="cuda"))
    noise = (torch.from_numpy(g["density_draws"]).cuda() - inp_s["density"][None]) / std[None]
    monkeypatch.setenv("UB_LAPLACE_TORCH_DRAWS", "1")
    monkeypatch.setattr(torch, "randn", lambda *a, **k: noise.reshape(100, 16, 48, 1))
    model = _laplace_standin(stubs, inp_s, dv, lv_s)
    out = model.get_outputs_unc(stubs.flat_ray_bundle(16, device="cuda"), is_inference=True, use_deterministic_density=False)
    monkeypatch.undo()
    _depth_close(out["depth"], g["samp_depth"], max_flips=1)
    for k in ("rgb", "rgb_std"):
        _close(out[k], g[f"samp_{k}"], k)
    for k in ("accumulation", "expected_depth"):       # the draws are reconstructed from (draw - mu) / std: 1-ulp noise
        torch.testing.assert_close(out[k].cpu(), torch.from_numpy(g[f"samp_{k}"]), rtol=2e-5, atol=2e-6)
    # default mode: in-kernel Philox draws -> statistical agreement with the reference's Monte-Carlo estimate
    out = model.get_outputs_unc(stubs.flat_ray_bundle(16, device="cuda"), is_inference=True, use_deterministic_density=False)
    assert float((out["accumulation"].cpu() - torch.from_numpy(g["samp_accumulation"])).abs().max()) < 0.05


@pytest.mark.parametrize("head,out_dim,act", [("rgb", 3, torch.nn.Sigmoid()), ("density", 1, torch.exp)])
def test_sample_laplace_replacement(stubs, monkeypatch, head, out_dim, act):
    """``NerfactoLaplaceField.sample_laplace`` replacement vs the reference's own loop on the same ``randn`` draws."""
    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    g = z("ref_laplace.npz")
    lap = synthetic.laplace_head(257, 64, out_dim, 100, seed=5 + out_dim)
    lin = torch.nn.Linear(64, out_dim).cuda()
    with torch.no_grad():
        lin.weight.copy_(lap["mu_q"][:64 * out_dim].view(out_dim, 64))
        lin.bias.copy_(lap["mu_q"][64 * out_dim:])
    draws = torch.from_numpy(g[f"{head}_randn"]).cuda()
    monkeypatch.setattr(torch, "randn", lambda *a, **k: draws)
    mu, s2 = plug.laplace_sample_laplace(None, lin, act, lap["ggn"].cuda(), lap["x"].cuda().view(257, 1, 64), 100, 1.0, 1e-9)
    monkeypatch.undo()
    assert mu.shape == (257, 1, out_dim) and s2.shape == (257, 1, out_dim)
    want_mu, want_s2 = torch.from_numpy(g[f"{head}_mu"]), torch.from_numpy(g[f"{head}_sigma2"])
    torch.testing.assert_close(mu.cpu().view(257, out_dim), want_mu, rtol=1e-5, atol=1e-6)
    mu2 = want_s2 + want_mu ** 2                       # sigma2 = E[y^2] - E[y]^2 cancels: floor at 1e-5 of E[y^2]
    assert bool(((s2.cpu().view(257, out_dim) - want_s2).abs() <= 1e-5 * mu2.abs() + 1e-5 * want_s2.abs() + 1e-7).all())


def test_sample_laplace_falls_back_for_unsupported_modules(stubs):
    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    called = {}
    fake_self = types.SimpleNamespace(_ub_reference_sample_laplace=lambda **kw: called.setdefault("kw", kw) or (1, 2))
    mlp = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.ReLU()).cuda()
    plug.laplace_sample_laplace(fake_self, mlp, torch.nn.Sigmoid(), torch.ones(80).cuda(), torch.ones(4, 8).cuda(), 10, 1.0)
    assert called["kw"]["n_samples"] == 10 and called["kw"]["module"] is mlp


def test_mcdropout_replacement_and_subclass_does_not_recurse(stubs):
    from nerfstudio.models.nerfacto import NerfactoModel, NerfactoModelConfig

    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    g = z("ref_reduce.npz")
    renders = [_cuda(o) for o in synthetic.member_renders(10, 9, 11, seed=5)]
    replay = stubs.ReplayModel(renders)

    class Parent(NerfactoModel):
        def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle):
            return replay.get_outputs_for_camera_ray_bundle(camera_ray_bundle)

    class MC(Parent):
        pass

    MC.get_outputs_for_camera_ray_bundle = plug.make_mcdropout_get_outputs(MC.__mro__[1].get_outputs_for_camera_ray_bundle)

    class UserSubclass(MC):                 # round-1's super(type(self), self) recursed forever here
        pass

    cfg = NerfactoModelConfig()
    cfg.mc_samples = 10
    model = UserSubclass(cfg).cuda()
    model.drop = torch.nn.Dropout(0.5)
    model.eval()
    out = model.get_outputs_for_camera_ray_bundle(None)
    assert replay.calls == 10 and not model.training and not model.drop.training
    assert list(out.keys()) == list(g["mcd_keys"])
    for k, v in out.items():
        torch.testing.assert_close(v.cpu(), torch.from_numpy(g[f"mcd_{k}"]), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("tag,k,seed,pred_std", [("ensB", 5, 3, False), ("ensA", 3, 4, True)])
def test_ensemble_replacement(stubs, tag, k, seed, pred_std):
    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    g = z("ref_reduce.npz")
    renders = [_cuda(o) for o in synthetic.member_renders(k, 9, 11, seed=seed, with_pred_std=pred_std)]
    pipeline_self = types.SimpleNamespace(models=[stubs.ReplayModel([dict(o, note="not a tensor")]) for o in renders])
    out = plug.ensemble_get_outputs(pipeline_self, None, obb_box=None)
    assert list(out.keys()) == list(g[f"{tag}_keys"])
    for kk, v in out.items():
        torch.testing.assert_close(v.cpu(), torch.from_numpy(g[f"{tag}_{kk}"]), rtol=1e-5, atol=1e-7)


def _splat_standin(gauss, log_unc, background):
    from nerfstudio.models.splatfacto import SplatfactoModel, SplatfactoModelConfig

    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    @dataclass
    class Cfg(SplatfactoModelConfig):
        beta_min: float = 0.01

    class StandIn(SplatfactoModel):
        get_outputs = plug.active_splatfacto_get_outputs
        log_uncertainties = property(lambda self: self.gauss_params["log_uncertainties"])

        def populate_modules(self):
            super().populate_modules()
            self.activation_uncertainty = torch.nn.Softplus()

    model = StandIn(Cfg(sh_degree=3), seed_gaussians=gauss)
    model.gauss_params["log_uncertainties"] = torch.nn.Parameter(log_unc.clone())
    model.background_color = torch.tensor(background)
    return model.cuda().eval()


def test_active_splatfacto_get_outputs(stubs):
    """Camera -> projection -> SH colours -> binning -> fused passes, against the reference's ``get_outputs(camera)``
    (run over the gsplat stand-in): keys, shapes, early returns and the images."""
    from nerfstudio.cameras.cameras import Cameras

    g = z("ref_splat.npz")
    H, W, G, seed, deg = (int(v) for v in g["meta"])
    sc = synthetic.gaussians_3d(G, H, W, seed=seed, sh_degree=deg)
    gauss = {"means": sc["means"], "scales": torch.log(sc["scales"]), "quats": sc["quats"],
             "features_dc": sc["sh_coeffs"][:, 0, :], "features_rest": sc["sh_coeffs"][:, 1:, :],
             "opacities": torch.logit(sc["opacities"])}
    model = _splat_standin(gauss, torch.from_numpy(g["log_unc"]), (0.1, 0.2, 0.3))
    cam = Cameras(torch.from_numpy(g["c2w"]).cuda(), sc["fx"], sc["fy"], sc["cx"], sc["cy"], W, H)
    assert model.get_outputs("not a camera") == {}
    with torch.no_grad():
        out = model.get_outputs(cam)
    assert list(out.keys()) == ["rgb", "depth", "accumulation", "background", "uncertainty", "rgb_var", "rgb_std",
                                "depth_var", "depth_std"]
    assert model.last_size == (H, W) and model.xys.shape == (G, 2) and model.radii.shape == (G,)
    torch.testing.assert_close(model.xys.cpu(), torch.from_numpy(g["out__xys"]), rtol=1e-5, atol=1e-3)
    for k in ("rgb", "accumulation", "uncertainty", "rgb_std", "rgb_var", "depth"):
        got, want = out[k].cpu(), torch.from_numpy(g[f"out_{k}"])
        bad = ~torch.isclose(got, want, rtol=2e-4, atol=2e-5)
        assert float(bad.float().mean()) <= 2e-3, (k, float(bad.float().mean()))
    # nothing visible -> the parent's empty outputs, like :239-240
    far = dict(gauss, means=gauss["means"] + torch.tensor([0.0, 0.0, -1e4]))
    empty = _splat_standin(far, torch.from_numpy(g["log_unc"]), (0.1, 0.2, 0.3)).get_outputs(cam)
    assert list(empty.keys()) == ["rgb", "depth", "accumulation", "background"] and float(empty["accumulation"].sum()) == 0.0


def test_patched_metrics_entry_points_accept_the_reference_call_forms(stubs):
    """``_ause`` is what replaces ``nerfuncertainty.metrics.ause``: the scorer calls it with flat CUDA (or CPU)
    float tensors (eval_uncertainty.py:336-358); ``auce`` with host numpy arrays (:376-378)."""
    from oracle import metrics as om
    from uncertainty_nerf_gs_b200 import metrics
    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    p, s, gt = synthetic.scoring_image(40, 50, seed=1)
    pro = om.rgb_metric_prologue(p, gt, s)
    r, e, v, a = plug._ause(pro["var"], pro["squared_error"], "rmse")          # CPU tensors in, like a CPU-resident caller
    r0, e0, v0, a0 = om.ause(pro["var"], pro["squared_error"], "rmse")
    np.testing.assert_allclose(v, v0, rtol=1e-5)
    np.testing.assert_allclose(a, a0, rtol=1e-5, atol=1e-9)
    std3 = pro["var"].sqrt().unsqueeze(-1).repeat(1, 3).numpy()
    d = metrics.auce(p.reshape(-1, 3).numpy(), std3, gt.reshape(-1, 3).numpy())
    d0 = om.auce(p.reshape(-1, 3).numpy(), std3, gt.reshape(-1, 3).numpy())
    assert np.array_equal(d["coverage_values"], d0["coverage_values"])
