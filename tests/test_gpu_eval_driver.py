"""GPU: the batched eval driver (score -> records -> reference aggregation -> metrics.json) against the
oracle running the reference's per-view loop semantics (eval_uncertainty.py:896-1077)."""
import json

import numpy as np
import pytest
import torch

from oracle import metrics as om
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_driver_matches_reference_aggregation(built_library, tmp_path):
    from uncertainty_nerf_gs_b200 import eval_driver, pipeline

    imgs = [synthetic.scoring_image(48, 64, seed=i) for i in range(5)]
    views = [({"rgb": p.cuda(), "rgb_std": s.cuda()}, g.cuda()) for p, s, g in imgs]
    records = eval_driver.score_views(views, batch_size=2)
    results = eval_driver.average_uncertainty_metrics(records)

    per_image, curves = [], {k: [] for k in ("err_var_mse", "coverage_values")}
    for p, s, g in imgs:
        d = om.unc_metrics_rgb(p, g, s)
        per_image.append(om.per_image_rgb_scalars(d))
        for k in curves:
            curves[k].append(np.asarray(d[k], dtype=np.float64))
    ref_scalars = om.aggregate_scalars(per_image)
    for k, v in ref_scalars.items():
        np.testing.assert_allclose(results[k], v, rtol=1e-5, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(results["err_var_mse"], om.aggregate_curves(curves["err_var_mse"]), rtol=1e-5)
    assert np.array_equal(results["coverage_values"], om.aggregate_curves(curves["coverage_values"]))
    assert "num_rays_per_sec" in results and "fps" in results

    out = tmp_path / "run" / "metrics.json"
    eval_driver.write_metrics_json(out, "exp", "active-nerfacto", "ckpt/step-000029999.ckpt", results)
    info = json.loads(out.read_text())
    assert list(info.keys()) == ["experiment_name", "method_name", "checkpoint", "results"]
    for k in ("rgb_ause_mse", "rgb_ause_mae", "rgb_ause_rmse", "rgb_mse", "rgb_rmse", "rgb_nll", "rgb_avg_var",
              "rgb_auc_abs_error", "rgb_auc_length", "rgb_auc_neg_error", "num_rays_per_sec", "fps"):
        assert isinstance(info["results"][k], float)
    assert "psnr" not in info["results"] and "depth_nll" not in info["results"]     # groups that were not evaluated
    eval_driver.save_curves(tmp_path / "plots", results)
    assert np.load(tmp_path / "plots" / "auce_rgb_empirical_coverage.npy").shape == (99,)


def test_driver_equals_the_reference_test_set_loop(built_library, tmp_path):
    """``tests/golden/ref_scoring.npz`` holds what the REFERENCE'S ``get_average_uncertainty_metrics`` returned
    (executed unmodified, tests/golden/make_golden.py) for 3 views with rgb + depth scoring and model-layer image
    metrics, and the ``auce_*.npy`` files its ``plot_auce_curves`` wrote: same keys in the same order, scalars to
    1e-5, coverage curves (integer counts / n) exactly."""
    import os

    from uncertainty_nerf_gs_b200 import eval_driver, pipeline

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_scoring.npz"))
    views = []
    for i in range(3):
        pr, st, gi = synthetic.scoring_image(30, 40, seed=20 + i)
        gen = torch.Generator().manual_seed(30 + i)
        di = torch.rand(30, 40, 1, generator=gen) * 4
        dsi = torch.rand(30, 40, 1, generator=gen) * 0.3 + 0.01
        views.append(({"rgb": pr.cuda(), "rgb_std": st.cuda(), "depth": di.cuda(), "depth_std": dsi.cuda()}, gi.cuda()))
    dgt = [torch.from_numpy(a).cuda() for a in g["set_depth_gt"]]
    scale = float(np.loadtxt(_scale_file(tmp_path, 0.9), delimiter=","))
    records = eval_driver.score_views(views, batch_size=2, depth_gt=dgt, depth_scale=scale, min_depth_std_for_nll=2.0,
                                      image_metrics_fn=lambda o, t: {"psnr": 20.0, "ssim": 0.9, "lpips": 0.1})
    results = eval_driver.average_uncertainty_metrics(records)
    keys = [k for k in results if isinstance(results[k], float)]
    assert keys == list(g["set_keys"])                                     # metrics.json order
    for k, want in zip(g["set_keys"], g["set_values"]):
        if k in ("num_rays_per_sec", "fps"):
            continue
        np.testing.assert_allclose(results[k], want, rtol=1e-5, atol=1e-9, err_msg=str(k))
    for output in ("rgb", "depth"):
        written = eval_driver.save_curves(tmp_path / "plots", results, output)
        assert len(written) == 6
        for path in written:
            name = os.path.basename(path)[:-4]
            got, want = np.load(path), g["npy_" + name]
            if name.endswith("empirical_coverage"):
                assert np.array_equal(got, want), name
            else:
                np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12, err_msg=name)
    out = tmp_path / "metrics.json"
    eval_driver.write_metrics_json(out, "exp", "nerfacto-laplace", "ckpt", results)
    assert list(json.loads(out.read_text())["results"].keys()) == list(g["set_keys"])


def _scale_file(tmp_path, a):
    p = tmp_path / "scale_parameters.txt"
    np.savetxt(str(p), np.array([a]), delimiter=",")                   # the text round trip of eval_uncertainty.py:432
    return str(p)


def test_view_stream_equals_per_view_evaluation(built_library):
    """ViewStream (batched scoring on a second stream, delayed read-back) must return, in view order, exactly what
    evaluating every view on its own returns."""
    from uncertainty_nerf_gs_b200 import pipeline, synthetic

    h, w, S, M = 24, 40, 48, 3
    views = []
    for v in range(7):
        members = [synthetic.ray_samples(h * w, S, seed=100 * v + i, device="cuda") for i in range(M)]
        _, _, gt = synthetic.scoring_image(h, w, seed=v, device="cuda")
        views.append((members, gt))
    want = [pipeline.evaluate_view(m, g, h, w, rays_per_chunk=256) for m, g in views]
    vs = pipeline.ViewStream(h, w, rays_per_chunk=256, score_batch=3)
    got = []
    for m, g in views:
        got += vs.push(m, g)
    assert len(got) == 3                       # the first batch of three is read back when the second is enqueued
    got += vs.flush()
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert list(a.keys()) == list(b.keys())
        for k in b:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True), k
