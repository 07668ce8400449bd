"""Oracle: per-pixel mean / variance across K MC-dropout passes or M ensemble members.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pure torch-CPU restatement of
``nerfuncertainty/models/mcdropout/mcdropout_models.py:121-126`` and
``nerfuncertainty/models/ensemble/ensemble_pipeline.py:159-190``.  PINNED: bit-equal to the reference's own
methods executed unmodified (``tests/test_oracle_pinned.py::test_live_ensemble_reduce / test_live_mcdropout_reduce``
through ``oracle/ref_exec.py``) and to ``tests/golden/ref_reduce.npz`` produced by them.
"""
from __future__ import annotations

from typing import Dict, List

import torch

Tensor = torch.Tensor

STD_KEYS = ("rgb", "depth", "expected_depth")


def mcdropout_reduce(outputs_list: List[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    """``mcdropout_models.py:121-126``: every key -> mean over passes; for rgb / depth /
    expected_depth additionally ``<k>_std = std over passes (unbiased) averaged over channels``,
    inserted right after ``k``."""
    out: Dict[str, Tensor] = {}
    for k in outputs_list[0].keys():
        stacked = torch.stack([o[k] for o in outputs_list], dim=0)
        out[k] = stacked.mean(dim=0)
        if k in STD_KEYS:
            out[k + "_std"] = stacked.std(dim=0).mean(dim=-1)[..., None]
    return out


def ensemble_reduce(outputs_list: List[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    """``ensemble_pipeline.py:159-190``.

    Branch A (members expose ``rgb_std`` and ``depth_std``): aleatoric = mean member variance,
    epistemic = unbiased variance of member means, both channel-averaged; combined var/std.
    Branch B: sample std of member means.  The loop runs over the members' keys in insertion
    order and unconditionally stores the member mean under ``k`` -- so in branch A the combined
    ``rgb_var`` / ``rgb_std`` / ``depth_var`` / ``depth_std`` written while ``k`` is ``rgb`` /
    ``depth`` are later overwritten by the plain member means (the reference's order quirk)."""
    first = outputs_list[0]
    has_pred_std = "rgb_std" in first.keys() and "depth_std" in first.keys()
    out: Dict[str, Tensor] = {}
    for k in first.keys():
        stacked = torch.stack([o[k] for o in outputs_list], dim=0)
        out[k] = stacked.mean(dim=0)
        if has_pred_std:
            if k in ("rgb", "depth"):
                alea = torch.stack([o[k + "_var"] for o in outputs_list], dim=0)
                out[k + "_var_alea"] = alea.mean(dim=0).mean(dim=-1).unsqueeze(-1)
                out[k + "_var_epi"] = stacked.var(dim=0).mean(dim=-1).unsqueeze(-1)
                out[k + "_var"] = out[k + "_var_epi"] + out[k + "_var_alea"]
                out[k + "_std"] = out[k + "_var"].sqrt()
        elif k in STD_KEYS:
            out[k + "_std"] = stacked.std(dim=0).mean(dim=-1).unsqueeze(-1)
    return out
