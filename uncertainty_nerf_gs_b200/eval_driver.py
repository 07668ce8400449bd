"""Batched evaluation driver: the reference's per-view loop (``get_average_uncertainty_metrics``,
scripts/eval_uncertainty.py:816-1079) and its ``metrics.json`` (:1159-1171) on top of the fused kernels.

The reference renders and scores one view at a time, synchronising ~600 times per image.  Here views are
scored in batches (one set of segmented launches per batch, one device->host copy), the per-view records
have a fixed layout (``pipeline.pack_record``) so that ranks can all-gather them, and the aggregation is
the reference's: float64 curve sums in view order divided by the number of images, float32
``torch.mean`` of the per-image python floats.  PSNR / SSIM / LPIPS and plots belong to the model /
presentation layer and are out of scope; pass ``image_metrics_fn`` to add such per-view scalars.
"""
from __future__ import annotations

import json
import time
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import metrics, pipeline

Tensor = torch.Tensor
View = Tuple[Dict[str, Tensor], Tensor]  # (model outputs with "rgb", "rgb_std" [H,W,C]; ground-truth rgb [H,W,3])


def score_views(views: Sequence[View], batch_size: int = 8, min_rgb_std_for_nll: float = 3e-2,
                first_view_id: int = 0, image_metrics_fn: Optional[Callable[[Dict[str, Tensor], Tensor], Dict[str, float]]] = None
                ) -> Tuple[np.ndarray, List[Dict[str, float]]]:
    """Score rendered views (same-sized views of a batch go through one segmented launch set).  Returns
    ``(records [V, RECORD_LEN] float64, per-view extra scalars)``."""
    records, extras = [], []
    i = 0
    while i < len(views):
        shape = views[i][0]["rgb"].shape
        j = i
        while j < len(views) and j - i < batch_size and views[j][0]["rgb"].shape == shape:
            j += 1
        t0 = time.time()
        pred = torch.stack([v[0]["rgb"] for v in views[i:j]])
        std = torch.stack([v[0]["rgb_std"] for v in views[i:j]])
        gt = torch.stack([v[1] for v in views[i:j]])
        ds = metrics.score_rgb_batch(pred, gt, std, min_rgb_std_for_nll)
        dt = (time.time() - t0) / (j - i)
        for k, d in enumerate(ds):
            d.update(metrics.per_image_rgb_scalars(d))
            records.append(pipeline.pack_record(first_view_id + i + k, d))
            h, w = shape[0], shape[1]
            extra = {"num_rays_per_sec": h * w / dt, "fps": 1.0 / dt}      # eval_uncertainty.py:948-952
            if image_metrics_fn is not None:
                extra.update(image_metrics_fn(views[i + k][0], views[i + k][1]))
            extras.append(extra)
        i = j
    return (np.stack(records) if records else np.zeros((0, pipeline.RECORD_LEN))), extras


def average_uncertainty_metrics(records: np.ndarray, extras: Sequence[Dict[str, float]]) -> Dict[str, object]:
    """Aggregate like eval_uncertainty.py:957-1077: curves / num_images, float32 mean of the scalars."""
    agg = pipeline.aggregate_records(records)
    for key in (extras[0].keys() if extras else ()):
        agg[key] = float(torch.mean(torch.tensor([e[key] for e in extras])))
    return agg


def write_metrics_json(path, experiment_name: str, method_name: str, checkpoint: str, results: Dict[str, object]) -> None:
    """The reference's output file (eval_uncertainty.py:1162-1169): scalars only under ``results``."""
    scalars = {k: v for k, v in results.items() if isinstance(v, (int, float))}
    info = {"experiment_name": experiment_name, "method_name": method_name, "checkpoint": str(checkpoint),
            "results": scalars}
    p = Path(path)
    p.parent.mkdir(parents=True, exist_ok=True)
    p.write_text(json.dumps(info, indent=2), "utf8")


def save_curves(out_dir, results: Dict[str, object], output: str = "rgb") -> None:
    """The ``.npy`` curve dumps the reference writes next to its plots (metrics/auce.py:130-141)."""
    d = Path(out_dir)
    d.mkdir(parents=True, exist_ok=True)
    for k in pipeline.CURVE_KEYS_99 + pipeline.CURVE_KEYS_100:
        if k in results:
            np.save(d / f"{output}_{k}.npy", np.asarray(results[k]))
