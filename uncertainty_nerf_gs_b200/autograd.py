"""Differentiable front end of the fused ray compositor (training mode).

``composite_rays_train`` returns the same per-ray outputs as the reference's ``get_outputs`` while
training (activenerfacto_model.py:94-127: no nan_to_num / clamp on the colours, median depth under
``no_grad``) plus the volume-rendering ``weights`` that the interlevel / distortion losses consume, and
back-propagates to ``density``, ``rgb`` and ``beta`` through ``ub_composite_rays_backward``.  The sampler's
``deltas / starts / ends`` receive no gradient (nerfstudio's proposal sampler detaches them).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Union

import torch

from . import _lib, ops

Tensor = torch.Tensor
_OUT_KEYS = ("rgb", "accumulation", "expected_depth", "rgb_var", "rgb_std", "depth_var", "depth_std", "weights")


class _CompositeRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, deltas, starts, ends, rgb, beta, background, rays_per_chunk):
        out = ops.composite_rays(density, deltas, starts, ends, rgb, beta, background=background, beta_mode="raw",
                                 rays_per_chunk=rays_per_chunk, eval_mode=False, return_weights=True)
        ctx.background = background
        ctx.rays_per_chunk = rays_per_chunk
        # outputs the loss does not use must arrive as None, not as zeros: a zero gradient pushed through the sqrt of
        # rgb_std / depth_std is 0 / 0 = NaN wherever the variance is 0 (torch never runs that backward either)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(density, deltas, starts, ends, rgb, beta, out["depth"], out["_workspace"])
        ctx.mark_non_differentiable(out["depth"])
        return tuple(out[k] for k in _OUT_KEYS) + (out["depth"],)

    @staticmethod
    def backward(ctx, g_rgb, g_acc, g_exp, g_var, g_std, g_dvar, g_dstd, g_w, _g_depth):
        density, deltas, starts, ends, rgb, beta, depth, workspace = ctx.saved_tensors
        lib = _lib.load()
        R, S = density.shape[0], density.shape[1]
        a = _lib.CompositeRaysBwdArgs()
        flat = lambda t: t.contiguous()
        keep = [flat(t) for t in (density, deltas, starts, ends, rgb, beta)]
        a.density, a.deltas, a.starts, a.ends, a.rgb, a.beta = (t.data_ptr() for t in keep)
        a.num_rays, a.num_samples = R, S
        bg = ctx.background
        if isinstance(bg, str):
            a.background_mode = {"last_sample": _lib.UB_BG_LAST_SAMPLE, "random": _lib.UB_BG_NONE,
                                 "none": _lib.UB_BG_NONE}[bg]
        else:
            a.background_mode = _lib.UB_BG_FIXED
            a.background_rgb = (C.c_float * 3)(*[float(v) for v in bg])
        a.rays_per_chunk = int(ctx.rays_per_chunk) if ctx.rays_per_chunk else 0
        a.depth, a.chunk_workspace = depth.data_ptr(), workspace.data_ptr()
        grads = [None if g is None else g.contiguous().float() for g in (g_rgb, g_acc, g_exp, g_var, g_std, g_dvar, g_dstd, g_w)]
        names = ("g_rgb", "g_accumulation", "g_expected_depth", "g_rgb_var", "g_rgb_std", "g_depth_var",
                 "g_depth_std", "g_weights")
        for n, g in zip(names, grads):
            setattr(a, n, None if g is None else g.data_ptr())
        need = ctx.needs_input_grad
        dev = density.device
        gd = torch.empty_like(keep[0]) if need[0] else None
        gc = torch.empty_like(keep[4]) if need[4] else None
        gb = torch.empty_like(keep[5]) if need[5] else None
        a.out_g_density, a.out_g_rgb, a.out_g_beta = ops._ptr(gd), ops._ptr(gc), ops._ptr(gb)
        with ops._guard(dev):
            _lib.check(lib.ub_composite_rays_backward(C.byref(a), ops._stream()))
        ops._count(1)
        return gd, None, None, None, gc, gb, None, None


def composite_rays_train(density: Tensor, deltas: Tensor, starts: Tensor, ends: Tensor, rgb: Tensor, beta: Tensor,
                         background: Union[str, Sequence[float]] = "last_sample",
                         rays_per_chunk: Optional[int] = None) -> Dict[str, Tensor]:
    """Training-mode fused compositing with gradients.  Inputs ``[R, S, 1]`` / ``[R, S, 3]`` CUDA float32
    (S in {16, 32, 48, 64, 96}); returns the active-nerfacto training outputs incl. ``weights [R, S, 1]``.
    The reference's stability guard runs in training too (activenerfacto_model.py:104-106): if any ``beta`` is NaN,
    all NaNs become 0 -- the same ``nan_to_num`` op as the reference's, so its autograd semantics (no gradient to
    the replaced entries) carry over."""
    if torch.isnan(beta).any():
        beta = torch.nan_to_num(beta, 0.0)
    res = _CompositeRaysFn.apply(density, deltas, starts, ends, rgb, beta, background, rays_per_chunk)
    out = dict(zip(_OUT_KEYS + ("depth",), res))
    return {"rgb": out["rgb"], "accumulation": out["accumulation"], "depth": out["depth"],
            "expected_depth": out["expected_depth"], "density": density, "rgb_var": out["rgb_var"],
            "rgb_std": out["rgb_std"], "depth_var": out["depth_var"], "depth_std": out["depth_std"],
            "weights": out["weights"]}


class _CompositeTilesFn(torch.autograd.Function):
    """``ub_composite_tiles_planes`` / ``..._backward`` as one differentiable op.  ``args`` = the colour planes."""

    @staticmethod
    def forward(ctx, xys, conics, opacities, gaussian_ids, tile_bins, height, width, background, *planes):
        outs, alpha, _ = ops.composite_tiles_planes(xys, conics, opacities, planes, gaussian_ids, tile_bins,
                                                    height, width, background)
        ctx.height, ctx.width, ctx.background = height, width, background
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(xys, conics, opacities, gaussian_ids, tile_bins, *planes)
        return (*outs, alpha)

    @staticmethod
    def backward(ctx, *grads):
        xys, conics, opacities, gaussian_ids, tile_bins, *planes = ctx.saved_tensors
        v_outs, v_alpha = grads[:-1], grads[-1]
        want = list(ctx.needs_input_grad[8:])
        v_xys, v_conics, v_opac, v_pl = ops.composite_tiles_planes_backward(
            xys, conics, opacities, planes, gaussian_ids, tile_bins, ctx.height, ctx.width, ctx.background,
            v_outs, v_alpha, want)
        v_pl = [None if v is None else v.view_as(p) for v, p in zip(v_pl, planes)]
        return (v_xys, v_conics, v_opac.view_as(opacities), None, None, None, None, None, *v_pl)


def composite_tiles_train(xys: Tensor, conics: Tensor, opacities: Tensor, planes: Sequence[Tensor],
                          gaussian_ids: Tensor, tile_bins: Tensor, height: int, width: int,
                          background: Optional[Sequence[float]] = None):
    """Differentiable fused tile compositing: ``([H, W, c_p] per colour plane, alpha [H, W, 1])`` with gradients
    to ``xys``, ``conics``, ``opacities`` and every plane -- the role gsplat's ``rasterize_gaussians`` autograd
    function plays in the reference (activesplatfacto_model.py:260-301), all planes in one pass each way."""
    res = _CompositeTilesFn.apply(xys, conics, opacities, gaussian_ids, tile_bins, height, width,
                                  None if background is None else [float(v) for v in background], *planes)
    return list(res[:-1]), res[-1]
