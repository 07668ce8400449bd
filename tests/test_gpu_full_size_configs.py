"""GPU: the remaining BASELINE.json configurations at their full sizes, checked through size-independent
properties (the CPU oracle needs seconds per image) plus an oracle spot check:

* configs[2]: AUSE / AUCE / NLL over 200 test views of 800 x 800 in one batch of segmented launches;
* configs[3]: active-splatfacto compositing of 1 M pre-binned Gaussians at 1297 x 840.
"""
import numpy as np
import pytest
import torch

from oracle import metrics as om
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_auce_ause_over_200_views(built_library):
    from uncertainty_nerf_gs_b200.metrics import score_rgb_batch

    b, h, w = 200, 800, 800
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.rand(b, h, w, 3, generator=g, device="cuda")
    std = torch.clamp(0.1 * torch.rand(b, h, w, 1, generator=g, device="cuda"), min=0.03)
    gt = torch.clamp(pred + std * torch.randn(b, h, w, 3, generator=g, device="cuda"), 0.0, 1.0)
    outs = score_rgb_batch(pred, gt, std)
    assert len(outs) == b
    n = h * w
    for d in outs[::17]:
        cov = d["coverage_values"]
        assert cov.shape == (99,) and (np.diff(cov) <= 0).all() and 0.0 <= cov[-1] <= cov[0] <= 1.0
        assert np.allclose(np.rint(cov * 3 * n), cov * 3 * n, atol=1e-6)       # integer counts / (3 N)
        assert d["err_mse"].shape == (100,) and np.isfinite(d["ause_rmse"])
        assert abs(d["err_var_mse"][0] - d["err_mse"][0]) < 1e-6              # full-image mean is order independent
        assert d["auc_abs_error_values"] >= 0 and d["avg_var"] > 0
    # spot check two images against the oracle (the reference's per-image path)
    for i in (0, 199):
        ref = om.unc_metrics_rgb(pred[i].cpu(), gt[i].cpu(), std[i].cpu())
        assert np.array_equal(outs[i]["coverage_values"], ref["coverage_values"])
        np.testing.assert_allclose(outs[i]["ause_mae"], ref["ause_mae"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(outs[i]["err_var_rmse"], ref["err_var_rmse"], rtol=1e-5)
        np.testing.assert_allclose(outs[i]["nll_rgb"], ref["nll_rgb"], rtol=1e-5)


def test_splat_one_million_gaussians_full_view(built_library):
    from uncertainty_nerf_gs_b200 import binning
    from uncertainty_nerf_gs_b200.models.outputs import active_splatfacto_outputs

    h, w, g = 840, 1297, 1_000_000
    sc = synthetic.splat_scene(g, h, w, seed=0, device="cuda")
    ids, bins = binning.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)
    assert bins.shape == (82 * 53, 2)
    from oracle import splat as osp
    ids_ref, bins_ref = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], h, w)   # torch restatement, on the GPU
    assert torch.equal(ids, ids_ref) and torch.equal(bins, bins_ref)
    d = sc["depths"][ids.long()]
    lo, hi = bins[:, 0].long(), bins[:, 1].long()
    inner = torch.ones(ids.numel(), dtype=torch.bool, device="cuda")
    inner[lo[lo < ids.numel()]] = False                                       # first entry of every tile
    assert bool((d[1:][inner[1:]] >= d[:-1][inner[1:]]).all())                # depth-sorted inside every tile
    bg = torch.tensor([0.2, 0.4, 0.6], device="cuda")
    out = active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"], sc["betas"],
                                    ids, bins, h, w, bg)
    a = out["accumulation"]
    assert a.shape == (h, w, 1) and float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    assert float(out["rgb"].max()) <= 1.0 and float(out["rgb"].min()) >= 0.0
    covered = a[..., 0] > 0
    dmin, dmax = float(sc["depths"].min()), float(sc["depths"].max())
    dep = out["depth"][..., 0][covered]
    assert float(dep.min()) >= dmin * (1 - 1e-4) and float(dep.max()) <= dmax * (1 + 1e-4)   # convex combination
    assert float(out["uncertainty"].min()) >= 0.0 and bool(torch.isfinite(out["depth_std"]).all())
    assert torch.equal(out["rgb_var"], out["uncertainty"] ** 2)
    # translation property: shifting every colour by c shifts the composited image by c * alpha (+ background)
    from uncertainty_nerf_gs_b200 import ops
    col = sc["rgbs"]
    base, al = ops.composite_tiles(sc["xys"], sc["conics"], sc["opacities"], col, ids, bins, h, w, [0, 0, 0])
    shifted, _ = ops.composite_tiles(sc["xys"], sc["conics"], sc["opacities"], col + 0.25, ids, bins, h, w, [0, 0, 0])
    torch.testing.assert_close(shifted, base + 0.25 * al, rtol=1e-4, atol=1e-5)


def test_full_view_compositing_against_the_oracle(built_library):
    """configs[1] at its real size: all 34 eval chunks of one 1297 x 840 view through the CPU oracle (which
    tests/test_oracle_pinned.py pins bit-for-bit to the reference's own ``get_outputs`` + chunk loop)."""
    from oracle import compositing as oc
    from uncertainty_nerf_gs_b200.models.outputs import active_nerfacto_outputs

    h, w, S, chunk = 840, 1297, 48, 1 << 15
    R = h * w
    inp = synthetic.ray_samples(R, S, seed=42)
    keys = ("density", "deltas", "starts", "ends", "rgb", "beta")
    ref = oc.render_in_chunks(oc.active_nerfacto_outputs, chunk, *[inp[k] for k in keys])
    out = active_nerfacto_outputs(*[inp[k].cuda() for k in keys], rays_per_chunk=chunk, image_hw=(h, w))
    assert out["rgb"].shape == (h, w, 3)
    same_depth = torch.isclose(out["depth"].reshape(R, 1).cpu(), ref["depth"], rtol=1e-6, atol=0.0)[:, 0]
    assert int((~same_depth).sum()) <= 1e-4 * R                           # borderline median samples only
    atol = {"rgb": 2e-6, "accumulation": 2e-6, "expected_depth": 2e-6, "rgb_var": 1e-7, "rgb_std": 1e-6}
    for k, a in atol.items():
        torch.testing.assert_close(out[k].reshape(R, -1).cpu(), ref[k], rtol=1e-5, atol=a, equal_nan=True, msg=lambda m: f"{k}: {m}")
    for k in ("depth_var", "depth_std"):
        torch.testing.assert_close(out[k].reshape(R, 1).cpu()[same_depth], ref[k][same_depth], rtol=1e-5, atol=1e-6)


def test_full_size_k10_reduce_against_the_oracle(built_library):
    """configs[2]'s reduce at full image size: K = 10 MC-dropout passes of a 1297 x 840 view, every nerfacto key."""
    from oracle import reduce as orc
    from uncertainty_nerf_gs_b200.models.outputs import mcdropout_reduce

    passes = synthetic.member_renders(10, 840, 1297, seed=7)
    ref = orc.mcdropout_reduce(passes)
    out = mcdropout_reduce([{k: v.cuda() for k, v in p.items()} for p in passes])
    assert list(out.keys()) == list(ref.keys())
    for k, v in ref.items():
        torch.testing.assert_close(out[k].cpu(), v, rtol=1e-5, atol=1e-7, msg=lambda m: f"{k}: {m}")


@pytest.mark.parametrize("sort_path", [True, False])
def test_full_scorer_at_800x800_both_ause_paths(built_library, monkeypatch, sort_path):
    """configs[0]'s image: the whole scorer at 800 x 800 through the per-image segmented radix sort (the north-star's
    kernel (c), ``UB_AUSE_SORT=1``) and through the sort-free select, against the oracle (= the reference's
    ``get_unc_metrics_rgb``, pinned): coverage counts exactly, curves / scalars to 1e-5."""
    from uncertainty_nerf_gs_b200 import metrics

    monkeypatch.setenv("UB_AUSE_SORT", "1" if sort_path else "0")
    p, s, g = synthetic.scoring_image(800, 800, seed=5)
    ref = om.unc_metrics_rgb(p, g, s)
    d = metrics.score_rgb_batch(p.cuda(), g.cuda(), s.cuda())[0]
    assert np.array_equal(d["coverage_values"], ref["coverage_values"])
    for k in ("err_mae", "err_mse", "err_rmse", "err_var_mae", "err_var_mse", "err_var_rmse"):
        np.testing.assert_allclose(np.asarray(d[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64), rtol=1e-5, err_msg=k)
    for k in ("ause_mae", "ause_mse", "ause_rmse", "nll_rgb", "avg_var", "auc_abs_error_values", "auc_length_values"):
        np.testing.assert_allclose(d[k], ref[k], rtol=1e-5, atol=1e-9, err_msg=k)
