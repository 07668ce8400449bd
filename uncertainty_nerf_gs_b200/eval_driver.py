"""Batched evaluation driver: the reference's per-view loop (``get_average_uncertainty_metrics``,
scripts/eval_uncertainty.py:816-1079), its ``metrics.json`` (:1159-1171) and the ``auce_*.npy`` dumps of
``plot_auce_curves`` (metrics/auce.py:130-141) on top of the fused kernels.

The reference renders and scores one view at a time, synchronising ~600 times per image.  Here views are
scored in batches (one set of segmented launches per batch and modality, one device->host copy), the per-view
records have a fixed layout (``pipeline.pack_record``) so that ranks can all-gather them, and the aggregation is
the reference's: float64 curve sums in view order divided by the number of images, float32 ``torch.mean`` of
the per-image python floats.  PSNR / SSIM / LPIPS are computed by the model layer (``model.psnr`` ...,
eval_uncertainty.py:680-686; torchmetrics networks, out of scope): pass ``image_metrics_fn`` and their values
travel in the record and come out under the reference's keys; plots are presentation and out of scope.
"""
from __future__ import annotations

import json
import time
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import metrics, pipeline

Tensor = torch.Tensor
# (model outputs with "rgb", "rgb_std" [H,W,C] and, for depth scoring, "depth", "depth_std" [H,W,1];
#  ground-truth rgb [H,W,3]); splat models also carry "background" (see ``composite_gt_with_background``)
View = Tuple[Dict[str, Tensor], Tensor]


def composite_gt_with_background(rgb_gt: Tensor, background: Tensor) -> Tensor:
    """``model.composite_with_background(model.get_gt_img(image), outputs["background"])`` of the splat models
    (eval_uncertainty.py:320-322; nerfstudio splatfacto): uint8 -> float, RGBA blended over the background."""
    if rgb_gt.dtype == torch.uint8:
        rgb_gt = rgb_gt.float() / 255.0
    if rgb_gt.shape[-1] == 4:
        alpha = rgb_gt[..., -1:].expand(*rgb_gt.shape[:-1], 3)
        return alpha * rgb_gt[..., :3] + (1 - alpha) * background.to(rgb_gt.device)
    return rgb_gt


def score_views(views: Sequence[View], batch_size: int = 8, min_rgb_std_for_nll: float = 3e-2,
                first_view_id: int = 0,
                image_metrics_fn: Optional[Callable[[Dict[str, Tensor], Tensor], Dict[str, float]]] = None,
                eval_rgb: bool = True, depth_gt: Optional[Sequence[Tensor]] = None,
                depth_scale: float = 1.0, min_depth_std_for_nll: float = 1.0) -> np.ndarray:
    """Score rendered views (same-sized views of a batch go through one segmented launch set per modality).
    ``depth_gt[i]`` (``[H, W]``, 0 = invalid) switches the depth modality on (``eval_depth``, the side inputs of
    eval_uncertainty.py:432-436 already loaded; ``depth_scale`` = ``scale_parameters.txt``).  Returns the
    ``[V, RECORD_LEN]`` float64 records."""
    records = []
    i = 0
    while i < len(views):
        shape = views[i][0]["rgb"].shape
        j = i
        while j < len(views) and j - i < batch_size and views[j][0]["rgb"].shape == shape:
            j += 1
        t0 = time.time()
        rgb_rows: List[Optional[Dict[str, object]]] = [None] * (j - i)
        depth_rows: List[Optional[Dict[str, object]]] = [None] * (j - i)
        if eval_rgb:
            pred = torch.stack([v[0]["rgb"] for v in views[i:j]])
            std = torch.stack([v[0]["rgb_std"] for v in views[i:j]])
            gt = torch.stack([composite_gt_with_background(v[1], v[0]["background"]) if "background" in v[0] else v[1]
                              for v in views[i:j]])
            rgb_rows = metrics.score_rgb_batch(pred, gt, std, min_rgb_std_for_nll)
            for d in rgb_rows:
                d.update(metrics.per_image_rgb_scalars(d))
        if depth_gt is not None:
            dp = torch.stack([v[0]["depth"] for v in views[i:j]])
            ds = torch.stack([v[0]["depth_std"] for v in views[i:j]])
            dg = torch.stack([depth_gt[k] for k in range(i, j)])
            depth_rows = metrics.score_depth_batch(dp, ds, dg, [depth_scale] * (j - i), min_depth_std_for_nll)
            for d in depth_rows:
                d.update(metrics.per_image_depth_scalars(d))
        dt = (time.time() - t0) / (j - i)
        h, w = shape[0], shape[1]
        for k in range(j - i):
            extra = {"num_rays_per_sec": h * w / dt, "fps": 1.0 / dt}               # eval_uncertainty.py:948-952
            if image_metrics_fn is not None:
                extra.update(image_metrics_fn(views[i + k][0], views[i + k][1]))
            records.append(pipeline.pack_record(first_view_id + i + k, rgb_rows[k], depth_rows[k], extra))
        i = j
    return np.stack(records) if records else np.zeros((0, pipeline.RECORD_LEN))


def average_uncertainty_metrics(records: np.ndarray) -> Dict[str, object]:
    """Aggregate like eval_uncertainty.py:957-1077: curves / num_images, float32 mean of the scalars, keys in
    the reference's ``metrics.json`` order."""
    return pipeline.aggregate_records(records)


def write_metrics_json(path, experiment_name: str, method_name: str, checkpoint: str, results: Dict[str, object]) -> None:
    """The reference's output file (eval_uncertainty.py:1162-1169): scalars only under ``results``."""
    scalars = {k: results[k] for k in pipeline.ALL_SCALAR_KEYS if k in results}
    info = {"experiment_name": experiment_name, "method_name": method_name, "checkpoint": str(checkpoint),
            "results": scalars}
    p = Path(path)
    p.parent.mkdir(parents=True, exist_ok=True)
    p.write_text(json.dumps(info, indent=2), "utf8")


_NPY_NAMES = {"coverage_values": "empirical_coverage", "avg_length_values": "avg_length",
              "coverage_error_values": "empirical_coverage_error",
              "abs_coverage_error_values": "empirical_coverage_absolute_error",
              "neg_coverage_error_values": "empirical_coverage_negative_error"}


def save_curves(out_dir, results: Dict[str, object], output: str = "rgb") -> List[str]:
    """The ``.npy`` dumps of ``plot_auce_curves`` (metrics/auce.py:130-141), same file names and contents:
    ``auce_{output}_alphas.npy`` and the five averaged calibration curves of the modality.  Returns the paths."""
    d = Path(out_dir)
    d.mkdir(parents=True, exist_ok=True)
    prefix = "" if output == "rgb" else output + "_"
    written = []
    if prefix + "coverage_values" not in results:
        return written
    path = d / f"auce_{output}_alphas.npy"
    np.save(path, list(np.arange(start=0.01, stop=1.0, step=0.01)))          # auce.py:60
    written.append(str(path))
    for key, name in _NPY_NAMES.items():
        path = d / f"auce_{output}_{name}.npy"
        np.save(path, np.asarray(results[prefix + key]))
        written.append(str(path))
    return written
