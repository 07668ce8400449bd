"""Output-dict assembly of the reference's models on top of the fused CUDA ops.

These functions are the part of each model's ``get_outputs*`` that lies downstream of the field /
projection (the hot path); they return dicts with the reference's keys *in the reference's insertion
order* (which matters for the ensemble overwrite quirk).  They work on plain tensors so that they can
be tested without nerfstudio; ``nerfstudio_plugin.py`` wraps them into nerfstudio ``Model`` subclasses.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from .. import ops

Tensor = torch.Tensor
Level = Tuple[Tensor, Tensor, Tensor]  # (weights, starts, ends) of one proposal level

STD_KEYS = ("rgb", "depth", "expected_depth")


def active_nerfacto_outputs(density: Tensor, deltas: Optional[Tensor], starts: Tensor, ends: Tensor, rgb: Tensor,
                            beta: Tensor, *, background="last_sample", rays_per_chunk: Optional[int] = None,
                            eval_mode: bool = True, proposal_levels: Sequence[Level] = (),
                            return_weights: bool = False, image_hw: Optional[Tuple[int, int]] = None
                            ) -> Dict[str, Tensor]:
    """``ActiveNerfactoModel.get_outputs`` downstream of ``field.forward``
    (reference activenerfacto_model.py:94-127, 150-151).  Inputs ``[R, S, C]``; outputs ``[R, C]``, or
    ``[H, W, C]`` with ``image_hw`` (the per-camera view the chunk loop of ``get_outputs_for_camera`` builds).
    ``deltas=None``: ``ends - starts``, see ``ops.composite_rays``."""
    o = ops.composite_rays(density, deltas, starts, ends, rgb, beta, background=background,
                           beta_mode="nan_guard", rays_per_chunk=rays_per_chunk, eval_mode=eval_mode,
                           return_weights=return_weights, image_hw=image_hw)
    out = {
        "rgb": o["rgb"],
        "accumulation": o["accumulation"],
        "depth": o["depth"],
        "expected_depth": o["expected_depth"],
        "density": density,
        "rgb_var": o["rgb_var"],
        "rgb_std": o["rgb_std"],
        "depth_var": o["depth_var"],
        "depth_std": o["depth_std"],
    }
    if return_weights:
        out["weights"] = o["weights"]
    for i, (w, s, e) in enumerate(proposal_levels):
        out[f"prop_depth_{i}"] = ops.render_weights(w, s, e, want=("depth",))["depth"]
    return out


def active_nerfacto_outputs_many(members: Sequence[Dict[str, Tensor]], *, background="last_sample",
                                 rays_per_chunk: Optional[int] = None, image_hw: Optional[Tuple[int, int]] = None
                                 ) -> List[Dict[str, Tensor]]:
    """``active_nerfacto_outputs`` (eval mode) for the M members of one view in one batched call
    (``ub_composite_rays_batch``).  ``members[i]`` holds ``density, deltas, starts, ends, rgb, beta`` (``deltas``
    absent or None: ``ends - starts``); every returned dict has the reference's keys in the reference's order."""
    res = ops.composite_rays_many([(m["density"], m.get("deltas"), m["starts"], m["ends"], m["rgb"], m["beta"])
                                   for m in members], background=background, beta_mode="nan_guard",
                                  rays_per_chunk=rays_per_chunk, eval_mode=True, image_hw=image_hw)
    return [{"rgb": o["rgb"], "accumulation": o["accumulation"], "depth": o["depth"],
             "expected_depth": o["expected_depth"], "density": m["density"], "rgb_var": o["rgb_var"],
             "rgb_std": o["rgb_std"], "depth_var": o["depth_var"], "depth_std": o["depth_std"]}
            for o, m in zip(res, members)]


def laplace_outputs_unc(density: Tensor, deltas: Tensor, starts: Tensor, ends: Tensor, rgb: Tensor,
                        rgb_var: Tensor, *, averaged_weights: Optional[Tensor] = None,
                        density_var: Optional[Tensor] = None, density_noise: Optional[Tensor] = None,
                        num_draws: int = 100, seed: int = 0, background="last_sample",
                        rays_per_chunk: Optional[int] = None, proposal_levels: Sequence[Level] = ()
                        ) -> Dict[str, Tensor]:
    """``NerfactoLaplaceModel.get_outputs_unc`` downstream of ``field.forward_unc``
    (reference laplace_model.py:471-530).  ``rgb`` / ``rgb_var`` are the last-layer Laplace moments
    (``laplace_ll_moments``).  With ``averaged_weights`` (mean weights of the sampled densities,
    :486-507) depth / depth_std / expected_depth / accumulation come from those, else from the
    deterministic weights (``use_deterministic_density``)."""
    o = ops.composite_rays(density, deltas, starts, ends, rgb, rgb_var, background=background, beta_mode="raw",
                           rays_per_chunk=rays_per_chunk, eval_mode=True)
    if averaged_weights is None and density_var is not None:
        # use_deterministic_density=False: mean weights of `num_draws` density draws (:486-507)
        averaged_weights = ops.average_sampled_weights(density, density_var, deltas, num_draws=num_draws,
                                                       noise=density_noise, seed=seed)
    if averaged_weights is not None:
        g = ops.render_weights(averaged_weights, starts, ends, rays_per_chunk=rays_per_chunk,
                               want=("accumulation", "depth", "expected_depth", "depth_std"))
    else:
        g = o
    out = {
        "rgb": o["rgb"],
        "rgb_std": o["rgb_std"],
        "accumulation": g["accumulation"],
        "depth": g["depth"],
        "depth_std": g["depth_std"],
        "expected_depth": g["expected_depth"],
    }
    for i, (w, s, e) in enumerate(proposal_levels):
        out[f"prop_depth_{i}"] = ops.render_weights(w, s, e, want=("depth",))["depth"]
    return out


def mcdropout_reduce(outputs_list: Sequence[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    """``NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle`` after the K stochastic renders
    (reference mcdropout_models.py:121-126): every key in one batched launch."""
    keys = list(outputs_list[0].keys())
    res = ops.reduce_many([([o[k] for o in outputs_list], "std" if k in STD_KEYS else None) for k in keys])
    out: Dict[str, Tensor] = {}
    for k, (mean, spread) in zip(keys, res):
        out[k] = mean
        if k in STD_KEYS:
            out[k + "_std"] = spread
    return out


def ensemble_reduce(outputs_list: Sequence[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    """``EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle`` after the M member renders
    (reference ensemble_pipeline.py:159-190), including the reference's overwrite order.  All member
    means / spreads come from one batched launch; only the two-term sums of branch A use torch."""
    first = outputs_list[0]
    keys = list(first.keys())
    has_pred_std = "rgb_std" in first.keys() and "depth_std" in first.keys()
    jobs = []
    for k in keys:
        if has_pred_std and k in ("rgb", "depth"):
            spread = "var"
        elif not has_pred_std and k in STD_KEYS:
            spread = "std"
        else:
            spread = None
        jobs.append(([o[k] for o in outputs_list], spread))
    res = dict(zip(keys, ops.reduce_many(jobs)))
    out: Dict[str, Tensor] = {}
    pos = {k: i for i, k in enumerate(keys)}
    for k in keys:
        mean, spread = res[k]
        out[k] = mean
        if has_pred_std and k in ("rgb", "depth"):
            alea = res[k + "_var"][0]                      # mean over members of the predicted variance
            # .mean(dim=-1, keepdim) of a one-channel image is the image itself (x / 1): skip the launch
            out[k + "_var_alea"] = alea if alea.shape[-1] == 1 else alea.mean(dim=-1).unsqueeze(-1)
            out[k + "_var_epi"] = spread
            # The reference goes on to store epi + alea and its square root under k_var / k_std -- and overwrites
            # both with the plain member means when its loop reaches those keys (the order quirk).  Entries that
            # a later key overwrites are not computed at all; the returned dict is the same.
            var_survives = pos.get(k + "_var", -1) < pos[k]
            std_survives = pos.get(k + "_std", -1) < pos[k]
            total = out[k + "_var_epi"] + out[k + "_var_alea"] if (var_survives or std_survives) else None
            # (an overwritten entry is inserted here with its final value, so the key order stays the reference's)
            out[k + "_var"] = total if var_survives else res[k + "_var"][0]
            out[k + "_std"] = total.sqrt() if std_survives else res[k + "_std"][0]
        elif not has_pred_std and k in STD_KEYS:
            out[k + "_std"] = spread
    return out


def active_splatfacto_outputs(xys: Tensor, depths: Tensor, conics: Tensor, opacities: Tensor, rgbs: Tensor,
                              betas: Tensor, gaussian_ids: Tensor, tile_bins: Tensor, height: int, width: int,
                              background: Tensor) -> Dict[str, Tensor]:
    """``ActiveSplatfactoModel.get_outputs`` downstream of projection and binning
    (reference activesplatfacto_model.py:260-367): one fused pass for rgb + beta + depth, the
    per-Gaussian squared depth residual, and a second pass for the depth variance."""
    g = xys.shape[0]
    bg = [float(v) for v in background.tolist()] + [0.0, 0.0]
    (rgb, uncertainty, depth_im), alpha, keys = ops.composite_tiles_planes(
        xys, conics, opacities, [rgbs.reshape(g, 3), betas.reshape(g, 1), depths.reshape(g, 1)], gaussian_ids,
        tile_bins, height, width, bg, want_max=True)
    ops.splat_normalize_(rgb, clamp_max_one=True)                       # rgb = clamp(rgb, max=1)         (:275)
    ops.splat_normalize_(depth_im, alpha, keys[4:5])                    # depth / alpha, else max(depth)   (:319)
    resid2 = ops.splat_depth_residual(xys, depths, depth_im)            # per-Gaussian squared residual    (:325-349)
    (depth_var,), _, vkeys = ops.composite_tiles_planes(xys, conics, opacities, [resid2], gaussian_ids, tile_bins,
                                                        height, width, [0.0], want_max=True)
    _, _, depth_std = ops.splat_normalize_(depth_var, alpha, vkeys[0:1], want_sqrt=True)   # / alpha, else max (:356); sqrt (:367)
    _, rgb_var, _ = ops.splat_normalize_(uncertainty, want_square=True)                     # uncertainty ** 2  (:364)
    return {
        "rgb": rgb,
        "depth": depth_im,
        "accumulation": alpha,
        "background": background,
        "uncertainty": uncertainty,
        "rgb_var": rgb_var,
        "rgb_std": uncertainty,
        "depth_var": depth_var,
        "depth_std": depth_std,
    }
