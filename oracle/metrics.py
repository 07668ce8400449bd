"""Oracle: AUSE / AUCE, the per-image metric prologue, NLL and the test-set aggregation.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``ause`` / ``auce`` are PINNED: ``tests/test_oracle_golden.py`` checks them against the
reference's own ``nerfuncertainty/metrics/ause.py:7-44`` and ``auce.py:10-57`` whenever
``/root/reference`` is mounted, and against golden vectors those functions produced
(``tests/golden/make_golden.py``).  The remaining functions restate
``nerfuncertainty/scripts/eval_uncertainty.py`` (cited per function) and are PINNED the same way:
``tests/test_oracle_pinned.py::test_live_rgb_scoring_and_nll / test_live_depth_scoring /
test_live_test_set_loop_and_npy_dumps`` execute ``get_unc_metrics_rgb``, ``get_unc_metrics_depth``,
``negative_gaussian_loglikelihood``, ``get_image_metrics_and_images_unc`` and ``get_average_uncertainty_metrics``
unmodified and demand bit equality (``tests/golden/ref_scoring.npz``).

Two contract decisions (hard parts 1 and 5):

* ranking uses ``torch.sort(stable=True)``.  The reference calls ``torch.sort`` with the
  default ``stable=False``, whose CPU result under ties is not reproducible; the stable
  permutation is what torch's CUDA radix path returns and what "stable tie-break" means.
  ``ause(..., stable=False)`` reproduces the reference's literal call for timing.
* AUCE follows NumPy >= 2 promotion: ``np.float64`` scalar * float32 array -> float64.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import scipy.stats
import torch

Tensor = torch.Tensor

if not hasattr(np, "trapz"):  # NumPy builds that dropped the alias
    np.trapz = np.trapezoid  # type: ignore[attr-defined]

N_RATIOS = 100
N_ALPHAS = 99


def ause_ratios() -> np.ndarray:
    """``ause.py:8``."""
    return np.linspace(0, 1, N_RATIOS, endpoint=False)


def ause_cut_counts(n: int) -> List[int]:
    """Prefix lengths ``int((1 - r) * n)`` exactly as ``ause.py:16,30`` evaluates them
    (float64 product, truncation)."""
    return [int((1 - r) * n) for r in ause_ratios()]


def _prefix_curve(sorted_err: Tensor, n: int, err_type: str) -> List[np.ndarray]:
    pts = []
    for c in ause_cut_counts(n):
        m = sorted_err[0:c].mean()
        if err_type == "rmse":
            m = torch.sqrt(m)
        elif err_type not in ("mae", "mse"):
            raise ValueError(err_type)
        pts.append(m.cpu().numpy())
    return pts


def ause(unc_vec: Tensor, err_vec: Tensor, err_type: str = "rmse", stable: bool = True
         ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.float64]:
    """Sparsification curves and their area -- ``nerfuncertainty/metrics/ause.py:7-44``.

    Returns ``(ratio_removed[100], oracle_curve[100], by_uncertainty_curve[100], ause)``
    with the reference's dtypes: the oracle curve is a list of 0-d float32 arrays divided by
    ``max_val``; the by-uncertainty curve is a float64 array.
    """
    ratios = ause_ratios()
    n = len(err_vec)
    err_sorted, _ = torch.sort(err_vec, stable=stable)
    oracle_pts = _prefix_curve(err_sorted, n, err_type)
    _, order = torch.sort(unc_vec, stable=stable)
    by_unc = np.zeros(len(ratios))
    for i, v in enumerate(_prefix_curve(err_vec[order], n, err_type)):
        by_unc[i] = v
    max_val = max(max(oracle_pts), max(by_unc))
    oracle_curve = np.array(oracle_pts / max_val)
    by_unc = np.array(by_unc / max_val)
    area = np.trapz(by_unc - oracle_curve, ratios)
    return ratios, oracle_curve, by_unc, area


def auce_alphas() -> List[np.float64]:
    """``auce.py:17`` -- 99 float64 values *with* their representation error."""
    return list(np.arange(start=0.01, stop=1.0, step=0.01))


def auce_z_values() -> np.ndarray:
    """``scipy.stats.norm.ppf(1 - alpha/2)`` per alpha (``auce.py:21-22``); strictly decreasing."""
    return np.array([scipy.stats.norm.ppf(1.0 - a / 2) for a in auce_alphas()], dtype=np.float64)


def auce(mean_values: np.ndarray, sigma_values: np.ndarray, target_values: np.ndarray) -> Dict[str, object]:
    """Calibration curves -- ``nerfuncertainty/metrics/auce.py:10-57`` (NumPy >= 2 semantics)."""
    n = float(np.prod(target_values.shape))
    alphas = auce_alphas()
    cov, length = [], []
    for a in alphas:
        z = scipy.stats.norm.ppf(1.0 - a / 2)
        lo = mean_values - z * sigma_values
        hi = mean_values + z * sigma_values
        inside = np.logical_and(target_values >= lo, target_values <= hi)
        cov.append(np.count_nonzero(inside) / n)
        length.append(np.mean(hi - lo))
    return auce_from_curves(np.array(cov), np.array(length))


def auce_from_curves(coverage: np.ndarray, avg_length: np.ndarray) -> Dict[str, object]:
    """Tail of ``auce.py:31-54``: errors, three trapezoid areas and the result dict."""
    alphas = auce_alphas()
    err = np.array(coverage) - (1.0 - np.array(alphas))
    abs_err = np.abs(err)
    neg_err = (np.abs(err) - err) / 2.0
    return {
        "coverage_values": np.array(coverage),
        "avg_length_values": np.array(avg_length),
        "coverage_error_values": np.array(err),
        "abs_coverage_error_values": abs_err,
        "neg_coverage_error_values": neg_err,
        "auc_abs_error_values": np.trapz(y=abs_err, x=alphas),
        "auc_length_values": np.trapz(y=list(avg_length), x=alphas),
        "auc_neg_error_values": np.trapz(y=neg_err, x=alphas),
    }


def negative_gaussian_loglikelihood(preds: Tensor, targets: Tensor, stds: Tensor, eps: float = 1e-6) -> Tensor:
    """``eval_uncertainty.py:404-412``."""
    s = stds.view(-1, 1)
    s = torch.maximum(s, torch.tensor([eps]))
    c = preds.shape[-1]
    dist = torch.distributions.Normal(loc=preds.view(-1, c), scale=s)
    return -dist.log_prob(targets.view(-1, c))


def rgb_metric_prologue(rgb_pred: Tensor, rgb_gt: Tensor, rgb_std: Tensor) -> Dict[str, Tensor]:
    """``eval_uncertainty.py:323-333``: per-pixel squared / absolute error summed over channels and
    the variance vector the AUSE ranks by."""
    se = torch.sum((rgb_pred - rgb_gt) ** 2, dim=-1).flatten()
    ae = torch.sum(torch.abs(rgb_pred - rgb_gt), dim=-1).flatten()
    var = (rgb_std ** 2).flatten()
    return {"squared_error": se, "absolute_error": ae, "var": var}


def unc_metrics_rgb(rgb_pred: Tensor, rgb_gt: Tensor, rgb_std: Tensor,
                    min_rgb_std_for_nll: float = 3e-2, stable: bool = True) -> Dict[str, object]:
    """``get_unc_metrics_rgb`` (``eval_uncertainty.py:306-402``) downstream of the background
    compositing of the ground truth; returns the scalar / curve entries of its ``dict_output``."""
    pro = rgb_metric_prologue(rgb_pred, rgb_gt, rgb_std)
    se, ae, var = pro["squared_error"], pro["absolute_error"], pro["var"]
    pred_flat = rgb_pred.reshape(-1, rgb_pred.shape[-1])
    gt_flat = rgb_gt.reshape(-1, rgb_gt.shape[-1])
    _, err_mae, err_var_mae, ause_mae = ause(var, ae, "mae", stable)
    _, err_mse, err_var_mse, ause_mse = ause(var, se, "mse", stable)
    _, err_rmse, err_var_rmse, ause_rmse = ause(var, se, "rmse", stable)
    nll = negative_gaussian_loglikelihood(pred_flat, gt_flat, rgb_std, eps=min_rgb_std_for_nll)
    std_flat = var.sqrt()
    if std_flat.dim() == 1:
        std_flat = std_flat.unsqueeze(-1).repeat(1, 3)
    out: Dict[str, object] = {
        "nll_rgb": torch.mean(nll).item(),
        "ause_mse": ause_mse, "ause_rmse": ause_rmse, "ause_mae": ause_mae,
        "err_mse": err_mse, "err_rmse": err_rmse, "err_mae": err_mae,
        "err_var_mse": err_var_mse, "err_var_rmse": err_var_rmse, "err_var_mae": err_var_mae,
        "mse": se.reshape(rgb_pred.shape[:-1]),
        "avg_var": var.mean().item(),
    }
    out.update(auce(pred_flat.numpy(), std_flat.numpy(), gt_flat.numpy()))
    return out


def resize_like_reference(img: Tensor, size) -> Tensor:
    """``F.resize(img.view(1, 1, H, W), size=size, antialias=None).squeeze(0, 1)`` (``eval_uncertainty.py:444-451``)."""
    h, w = img.shape[-2:]
    return torch.nn.functional.interpolate(img.reshape(1, 1, h, w), size=tuple(size), mode="bilinear",
                                           align_corners=False, antialias=False).squeeze(0).squeeze(0)


def unc_metrics_depth(depth: Tensor, depth_std: Tensor, depth_gt: Tensor, scale: float,
                      min_depth_std_for_nll: float = 1.0, stable: bool = True) -> Dict[str, object]:
    """``get_unc_metrics_depth`` (``eval_uncertainty.py:415-644``) downstream of file loading: the resize of
    renders whose shape differs from the ground truth's (``:442-452``; torchvision's ``F.resize`` of a float tensor =
    bilinear ``interpolate``, ``align_corners=False``, no antialiasing), scale, clamp to ``[1e-3, max gt]``, NLL on the
    full image, mask ``gt > 0``, errors, 3 x AUSE, AUCE."""
    depth = depth.squeeze(-1).clone()
    depth_std = depth_std.squeeze(-1).clone()
    if depth_gt.shape[-2:] != depth.shape[-2:]:
        depth = resize_like_reference(depth, depth_gt.shape[-2:])
    if depth_gt.shape[-2:] != depth_std.shape[-2:]:
        depth_std = resize_like_reference(depth_std, depth_gt.shape[-2:])
    min_d, max_d = 1e-3, depth_gt.max().float()
    depth = scale * depth
    depth_std = scale * depth_std
    clamped = depth.clone()
    clamped[clamped < min_d] = min_d
    clamped[clamped > max_d] = max_d
    nll_img = negative_gaussian_loglikelihood(clamped.unsqueeze(-1), depth_gt.unsqueeze(-1),
                                              depth_std.unsqueeze(-1), eps=min_depth_std_for_nll
                                              ).reshape(clamped.shape)
    mask = depth_gt > 0
    d, g, s = depth[mask], depth_gt[mask], depth_std[mask]
    d[d < min_d] = min_d
    d[d > max_d] = max_d
    se = (g - d) ** 2
    ae = abs(g - d)
    var = (s ** 2).flatten()
    _, err_mse, err_var_mse, ause_mse = ause(var, se.flatten(), "mse", stable)
    _, err_mae, err_var_mae, ause_mae = ause(var, ae.flatten(), "mae", stable)
    _, err_rmse, err_var_rmse, ause_rmse = ause(var, se.flatten(), "rmse", stable)
    out: Dict[str, object] = {
        "nll_depth": torch.mean(nll_img[mask]).item(),
        "ause_mse": ause_mse, "ause_rmse": ause_rmse, "ause_mae": ause_mae,
        "err_mse": err_mse, "err_rmse": err_rmse, "err_mae": err_mae,
        "err_var_mse": err_var_mse, "err_var_rmse": err_var_rmse, "err_var_mae": err_var_mae,
        "mse": se,
        "avg_var": var.mean().item(),
    }
    out.update(auce(d.flatten().numpy(), s.flatten().numpy(), g.flatten().numpy()))
    return out


RGB_SCALAR_KEYS = ("rgb_ause_mse", "rgb_ause_mae", "rgb_ause_rmse", "rgb_mse", "rgb_rmse", "rgb_nll",
                   "rgb_avg_var", "rgb_auc_abs_error", "rgb_auc_length", "rgb_auc_neg_error")


def per_image_rgb_scalars(d: Dict[str, object]) -> Dict[str, float]:
    """``eval_uncertainty.py:765-776``: the per-image scalar entries of ``metrics_dict``."""
    mse = float(d["mse"].mean().item())  # type: ignore[union-attr]
    return {
        "rgb_ause_mse": float(d["ause_mse"]), "rgb_ause_mae": float(d["ause_mae"]),
        "rgb_ause_rmse": float(d["ause_rmse"]), "rgb_mse": mse, "rgb_rmse": float(np.sqrt(mse)),
        "rgb_nll": float(d["nll_rgb"]), "rgb_avg_var": float(d["avg_var"]),
        "rgb_auc_abs_error": d["auc_abs_error_values"], "rgb_auc_length": d["auc_length_values"],
        "rgb_auc_neg_error": d["auc_neg_error_values"],
    }


def aggregate_scalars(per_image: Sequence[Dict[str, float]]) -> Dict[str, float]:
    """``eval_uncertainty.py:1070-1077``: float32 mean of the per-image python floats."""
    return {k: float(torch.mean(torch.tensor([m[k] for m in per_image]))) for k in per_image[0].keys()}


def aggregate_curves(per_image_curves: Sequence[np.ndarray]) -> np.ndarray:
    """``eval_uncertainty.py:920-946,957-1016``: float64 running sum in view order, then / num_images."""
    total = np.zeros(len(per_image_curves[0]))
    for c in per_image_curves:
        total += c
    return total / len(per_image_curves)
