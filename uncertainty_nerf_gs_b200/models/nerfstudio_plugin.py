"""nerfstudio-side glue: put the fused CUDA path behind the reference's plugin surface.

Nothing here imports nerfstudio (or the reference package ``nerfuncertainty``) at module import time --
neither is installable in the build container -- so the rest of the package stays testable without
them.  With both installed, ``patch_reference_models()`` swaps the hot-path arithmetic of the reference's
models for the ub200 kernels in place, keeping the method names, config classes, entry points
(reference pyproject.toml:18-22; ``uncertainty_nerf_gs_b200.models.method_configs`` re-exports them) and
output keys untouched.  Every function below is a plain ``def f(self, ...)`` with the reference method's
signature, so it can also be bound to a stand-in object (``tests/test_gpu_plugin.py`` does, on the GPU box):

* ``ActiveNerfactoModel.get_outputs``                 reference activenerfacto_model.py:83-152 (eval: fused
  compositor; training: the same kernel with its backward, ``autograd.composite_rays_train``)
* ``NerfactoLaplaceModel.get_outputs_unc``            reference laplace_model.py:456-556
* ``NerfactoLaplaceField.sample_laplace``             reference laplace_field.py:528-568 (``forward_unc`` /
  ``get_outputs`` stay the reference's, so the default-argument quirk of :516-520 -- the rgb head never sees the
  caller's ``prior_prec / n_samples / eps`` -- is preserved by construction)
* ``NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle``    reference mcdropout_models.py:94-131
* ``EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle`` reference ensemble_pipeline.py:144-191
* ``ActiveSplatfactoModel.get_outputs``               reference activesplatfacto_model.py:142-367 (eval; training keeps
  gsplat's autograd path: the projection / SH kernels here are forward-only)
* ``nerfuncertainty.metrics.ause / auce``             reference metrics/ause.py, auce.py
"""
from __future__ import annotations

import os
import warnings
from typing import Dict, List

import torch

from . import outputs as mo
from .. import metrics as ub_metrics

METHOD_NAMES = ("active-nerfacto", "active-splatfacto", "nerfacto-mcdropout", "nerfacto-laplace")
ENSEMBLE_METHOD_NAMES = ("nerfacto", "active-nerfacto", "splatfacto", "active-splatfacto")  # ensemble_utils.py:150-157
PATCHED_SURFACES = (
    "ActiveNerfactoModel.get_outputs",
    "NerfactoLaplaceModel.get_outputs_unc",
    "NerfactoLaplaceField.sample_laplace",
    "NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle",
    "EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle",
    "ActiveSplatfactoModel.get_outputs",
    "nerfuncertainty.metrics.ause",
    "nerfuncertainty.metrics.auce",
)


def _levels(weights_list, ray_samples_list, n):
    return [(weights_list[i], ray_samples_list[i].frustums.starts, ray_samples_list[i].frustums.ends)
            for i in range(n)]


def _flat(t, n_rays):
    return t.reshape(n_rays, -1, t.shape[-1]).float()


def _background_of(model):
    """The colour ``RGBRenderer`` would blend with: the global override if set (nerfstudio
    ``renderers.BACKGROUND_COLOR_OVERRIDE``), else the renderer's; names resolved through nerfstudio's colour table."""
    bg = model.renderer_rgb.background_color
    try:
        from nerfstudio.model_components import renderers

        if renderers.BACKGROUND_COLOR_OVERRIDE is not None:
            bg = renderers.BACKGROUND_COLOR_OVERRIDE
    except Exception:
        pass
    if isinstance(bg, str):
        if bg in ("last_sample", "random"):
            return bg
        named = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0), "red": (1.0, 0.0, 0.0), "green": (0.0, 1.0, 0.0),
                 "blue": (0.0, 0.0, 1.0)}
        if bg in named:
            return named[bg]
        from nerfstudio.utils.colors import get_color

        return tuple(float(v) for v in get_color(bg))
    return tuple(float(v) for v in (bg.tolist() if torch.is_tensor(bg) else bg))


# --------------------------------------------------------------------------------------------------------
def active_nerfacto_get_outputs(self, ray_bundle):
    """Replacement for ``ActiveNerfactoModel.get_outputs`` (activenerfacto_model.py:83-152).  One call = one eval
    chunk (or one training batch), so chunk-wide reductions span exactly the rays of this call."""
    if self.config.predict_normals or self.config.use_gradient_scaling:
        return self._ub_reference_get_outputs(ray_bundle)      # normals / gradient scaling: not on the fused path
    from nerfstudio.field_components.field_heads import FieldHeadNames

    ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
    field_outputs = self.field.forward(ray_samples, compute_normals=False)
    fr = ray_samples.frustums
    n_rays = fr.starts.shape[0]
    density = field_outputs[FieldHeadNames.DENSITY]
    args = (_flat(density, n_rays), _flat(ray_samples.deltas, n_rays), _flat(fr.starts, n_rays), _flat(fr.ends, n_rays),
            _flat(field_outputs[FieldHeadNames.RGB], n_rays), _flat(field_outputs["rgb_var"], n_rays))
    levels = [(_flat(w, n_rays), _flat(s, n_rays), _flat(e, n_rays))
              for w, s, e in _levels(weights_list, ray_samples_list, self.config.num_proposal_iterations)]
    if self.training:
        from ..autograd import composite_rays_train
        from .. import ops

        o = composite_rays_train(*args, background=_background_of(self))
        weights = o["weights"]
        weights_list.append(weights)
        ray_samples_list.append(ray_samples)
        out = {k: o[k] for k in ("rgb", "accumulation", "depth", "expected_depth")}
        out["density"] = density
        out.update({k: o[k] for k in ("rgb_var", "rgb_std", "depth_var", "depth_std")})
        out["weights_list"] = weights_list
        out["ray_samples_list"] = ray_samples_list
        with torch.no_grad():
            for i, (w, s, e) in enumerate(levels):
                out[f"prop_depth_{i}"] = ops.render_weights(w.detach(), s, e, want=("depth",))["depth"]
        return out
    if os.environ.get("UB_DERIVE_DELTAS", "0") == "1":
        # opt-in: RayBundle.get_ray_samples stores deltas = bin_ends - bin_starts and the same two tensors as the
        # frustum's starts / ends, so the compositor can take the difference itself and read one stream less
        # (bit-identical outputs, tests/test_gpu_composite.py).  Off by default: a sampler that fills
        # RaySamples.deltas any other way would be silently ignored.
        args = (args[0], None) + args[2:]
    out = mo.active_nerfacto_outputs(*args, background=_background_of(self), eval_mode=True, proposal_levels=levels)
    out["density"] = density
    return out


# --------------------------------------------------------------------------------------------------------
def _activation_name(activation):
    if isinstance(activation, torch.nn.Sigmoid) or activation is torch.sigmoid:
        return "sigmoid"
    if isinstance(activation, torch.nn.Identity):
        return "identity"
    owner = getattr(activation, "__self__", None)
    name = getattr(activation, "__name__", "") or getattr(owner, "__name__", "")
    if activation is torch.exp or name in ("trunc_exp", "_TruncExp", "exp") or getattr(owner, "__name__", "") == "_TruncExp":
        return "exp"            # nerfstudio's trunc_exp is exp in the forward direction
    return None


def laplace_sample_laplace(self, module, activation, diag_ggn, input, n_samples, prior_prec, eps=1e-9):
    """Replacement for ``NerfactoLaplaceField.sample_laplace`` (laplace_field.py:528-568): the posterior draws come
    from ``torch.randn`` on the global generator exactly as in the reference (same device + seed -> same draws); the
    loop of ``n_samples`` ``nn.Linear`` calls and the moment sums run as one fused kernel (tcgen05 for the 3-channel
    rgb head).  Anything that is not a CUDA ``nn.Linear`` with a supported activation takes the reference path."""
    from torch.nn.utils import parameters_to_vector

    from .. import ops

    act = _activation_name(activation)
    if not (isinstance(module, torch.nn.Linear) and module.bias is not None and act is not None and input.is_cuda):
        return self._ub_reference_sample_laplace(module=module, activation=activation, diag_ggn=diag_ggn, input=input,
                                                 n_samples=n_samples, prior_prec=prior_prec, eps=eps)
    mu_q = parameters_to_vector(module.parameters())
    n_params = len(mu_q)
    precision_matrix = (diag_ggn + prior_prec).to(mu_q.device)
    diag_covariance_matrix = 1 / torch.sqrt(precision_matrix + eps)
    samples = torch.randn(n_samples, n_params, device=mu_q.device)
    samples = samples * diag_covariance_matrix.view(1, n_params)
    samples_weights = mu_q.view(1, n_params) + samples
    out_dim = module.out_features
    with torch.no_grad():
        m = ops.laplace_ll_moments(input.reshape(-1, input.shape[-1]).float(), samples_weights.float(), out_dim, act)
    shape = (*input.shape[:-1], out_dim)
    return m["mean"].view(shape), m["sigma2"].view(shape)


def laplace_get_outputs_unc(self, ray_bundle, is_inference: bool = False, use_deterministic_density: bool = False,
                            prior_prec: float = 1.0, n_samples: int = 100, eps: float = 1e-9):
    """Replacement for ``NerfactoLaplaceModel.get_outputs_unc`` (laplace_model.py:456-556).  rgb / rgb_std from the
    deterministic weights; with sampled density the mean weights of 100 draws ``relu(N(density, sqrt(density_var)))``
    are accumulated in registers (the reference materialises ``[100, R, S, 1]``).  The draws come from an in-kernel
    Philox generator (statistical parity) unless ``UB_LAPLACE_TORCH_DRAWS=1``, which draws the standard normals with
    ``torch.randn`` on the global generator the way ``Normal.sample`` does (draw-for-draw parity, reference memory)."""
    if self.training or self.config.predict_normals or self.config.use_gradient_scaling:
        return self._ub_reference_get_outputs_unc(ray_bundle, is_inference, use_deterministic_density=use_deterministic_density,
                                                  prior_prec=prior_prec, n_samples=n_samples, eps=eps)
    from nerfstudio.field_components.field_heads import FieldHeadNames

    ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
    field_outputs = self.field.forward_unc(ray_samples, compute_normals=False, is_inference=is_inference,
                                           use_deterministic_density=use_deterministic_density, prior_prec=prior_prec,
                                           n_samples=n_samples, eps=eps)
    fr = ray_samples.frustums
    n_rays = fr.starts.shape[0]
    density = _flat(field_outputs[FieldHeadNames.DENSITY], n_rays)
    kw = {}
    if not use_deterministic_density:
        dvar = _flat(field_outputs["density_var"], n_rays)
        kw["density_var"] = dvar
        if os.environ.get("UB_LAPLACE_TORCH_DRAWS", "0") == "1":
            kw["density_noise"] = torch.randn(100, *density.shape, device=density.device)
        else:
            kw["seed"] = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())     # follows torch.manual_seed
    levels = [(_flat(w, n_rays), _flat(s, n_rays), _flat(e, n_rays))
              for w, s, e in _levels(weights_list, ray_samples_list, self.config.num_proposal_iterations)]
    return mo.laplace_outputs_unc(density, _flat(ray_samples.deltas, n_rays), _flat(fr.starts, n_rays),
                                  _flat(fr.ends, n_rays), _flat(field_outputs[FieldHeadNames.RGB], n_rays),
                                  _flat(field_outputs["rgb_var"], n_rays), background=_background_of(self),
                                  proposal_levels=levels, **kw)


# --------------------------------------------------------------------------------------------------------
def make_mcdropout_get_outputs(parent_method):
    """Replacement for ``NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle`` (mcdropout_models.py:94-131):
    K stochastic renders by the parent class (bound explicitly -- ``parent_method`` is the function the reference's
    zero-argument ``super()`` resolves to, so subclasses of the model do not recurse), one fused reduce."""

    @torch.no_grad()
    def mcdropout_get_outputs_for_camera_ray_bundle(self, camera_ray_bundle):
        def enable_dropout(mod):
            if isinstance(mod, torch.nn.Dropout):
                mod.train()

        train_status = self.training
        if train_status is False:
            self.apply(enable_dropout)
        outputs_list = [parent_method(self, camera_ray_bundle) for _ in range(self.config.mc_samples)]
        outputs = mo.mcdropout_reduce(outputs_list)
        if train_status is False:
            self.eval()
        return outputs

    return mcdropout_get_outputs_for_camera_ray_bundle


def ensemble_get_outputs(self, camera_ray_bundle, obb_box=None):
    """Replacement for ``EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle``
    (ensemble_pipeline.py:144-191)."""
    outputs_list = [m.get_outputs_for_camera(camera_ray_bundle, obb_box=obb_box) for m in self.models]
    tensors_only: List[Dict[str, torch.Tensor]] = [{k: v for k, v in o.items() if torch.is_tensor(v)}
                                                   for o in outputs_list]
    with torch.no_grad():
        return mo.ensemble_reduce(tensors_only)


# --------------------------------------------------------------------------------------------------------
def active_splatfacto_get_outputs(self, camera):
    """Replacement for ``ActiveSplatfactoModel.get_outputs`` in eval mode (activesplatfacto_model.py:142-367):
    same camera handling, cropping, early returns and attribute side effects (``self.xys``, ``self.radii``,
    ``self.last_size``); projection, SH colours, ONE tile binning (gsplat re-bins inside each of its four
    rasterisation calls) and the fused compositing passes run on the ub200 kernels."""
    if self.training:
        return self._ub_reference_get_outputs(camera)
    from nerfstudio.cameras.cameras import Cameras
    from nerfstudio.model_components import renderers

    from .. import binning

    if not isinstance(camera, Cameras):
        print("Called get_outputs with not a camera")
        return {}
    assert camera.shape[0] == 1, "Only one camera at a time"
    optimized_camera_to_world = self.camera_optimizer.apply_to_camera(camera)[0, ...]
    if renderers.BACKGROUND_COLOR_OVERRIDE is not None:
        background = renderers.BACKGROUND_COLOR_OVERRIDE.to(self.device)
    else:
        background = self.background_color.to(self.device)
    if self.crop_box is not None:
        crop_ids = self.crop_box.within(self.means).squeeze()
        if crop_ids.sum() == 0:
            return self.get_empty_outputs(int(camera.width.item()), int(camera.height.item()), background)
    else:
        crop_ids = None
    camera_downscale = self._get_downscale_factor()
    camera.rescale_output_resolution(1 / camera_downscale)
    R = optimized_camera_to_world[:3, :3]
    T = optimized_camera_to_world[:3, 3:4]
    R_edit = torch.diag(torch.tensor([1, -1, -1], device=self.device, dtype=R.dtype))
    R = R @ R_edit
    R_inv = R.T
    T_inv = -R_inv @ T
    viewmat = torch.eye(4, device=R.device, dtype=R.dtype)
    viewmat[:3, :3] = R_inv
    viewmat[:3, 3:4] = T_inv
    cx, cy = camera.cx.item(), camera.cy.item()
    W, H = int(camera.width.item()), int(camera.height.item())
    self.last_size = (H, W)
    pick = (lambda t: t[crop_ids]) if crop_ids is not None else (lambda t: t)
    opacities_crop, means_crop = pick(self.opacities), pick(self.means)
    features_dc_crop, features_rest_crop = pick(self.features_dc), pick(self.features_rest)
    scales_crop, quats_crop, log_unc_crop = pick(self.scales), pick(self.quats), pick(self.log_uncertainties)
    colors_crop = torch.cat((features_dc_crop[:, None, :], features_rest_crop), dim=1)
    with torch.no_grad():
        self.xys, depths, self.radii, conics, comp, num_tiles_hit, _cov3d = binning.project_gaussians(
            means_crop, torch.exp(scales_crop), 1, quats_crop / quats_crop.norm(dim=-1, keepdim=True),
            viewmat.squeeze()[:3, :], camera.fx.item(), camera.fy.item(), cx, cy, H, W, 16)
        camera.rescale_output_resolution(camera_downscale)
        if (self.radii).sum() == 0:
            return self.get_empty_outputs(W, H, background)
        if self.config.sh_degree > 0:
            viewdirs = means_crop.detach() - optimized_camera_to_world.detach()[:3, 3]
            n = min(self.step // self.config.sh_degree_interval, self.config.sh_degree)
            rgbs = torch.clamp(binning.spherical_harmonics(n, viewdirs, colors_crop) + 0.5, min=0.0)
        else:
            rgbs = torch.sigmoid(colors_crop[:, 0, :])
        if self.config.rasterize_mode == "antialiased":
            opacities = torch.sigmoid(opacities_crop) * comp[:, None]
        elif self.config.rasterize_mode == "classic":
            opacities = torch.sigmoid(opacities_crop)
        else:
            raise ValueError("Unknown rasterize_mode: %s", self.config.rasterize_mode)
        uncertainties = self.activation_uncertainty(log_unc_crop) + self.config.beta_min
        ids, bins = binning.bin_gaussians(self.xys, depths, self.radii, H, W)
        return mo.active_splatfacto_outputs(self.xys, depths, conics, opacities, rgbs, uncertainties, ids, bins, H, W,
                                            background)


# --------------------------------------------------------------------------------------------------------
def _ause(unc_vec, err_vec, err_type="rmse"):
    return ub_metrics.ause(unc_vec.cuda(), err_vec.cuda(), err_type)


def patch_reference_models() -> List[str]:
    """Swap the hot-path methods of the installed reference package for the ub200 ones.  Returns the list
    of patched qualified names (``PATCHED_SURFACES``); raises ImportError when nerfstudio / nerfuncertainty are
    missing.  The reference's own methods stay reachable as ``_ub_reference_*`` attributes (fallback for the
    configurations the fused path does not cover: predicted normals, gradient scaling, splat training)."""
    import sys

    import numpy as np

    import nerfuncertainty.metrics as ref_metrics
    from nerfuncertainty.models.activenerfacto.activenerfacto_model import ActiveNerfactoModel
    from nerfuncertainty.models.activesplatfacto.activesplatfacto_model import ActiveSplatfactoModel
    from nerfuncertainty.models.ensemble.ensemble_pipeline import EnsemblePipeline
    from nerfuncertainty.models.laplace.laplace_field import NerfactoLaplaceField
    from nerfuncertainty.models.laplace.laplace_model import NerfactoLaplaceModel
    from nerfuncertainty.models.mcdropout.mcdropout_models import NerfactoMCDropoutModel

    if int(np.__version__.split(".")[0]) < 2:
        warnings.warn("ub200 auce() evaluates the interval predicate in float64 (NumPy >= 2 promotion); under NumPy "
                      f"{np.__version__} the reference computes the bounds in float32, so coverage counts can differ "
                      "for elements that sit exactly on an interval boundary")

    def keep(cls, name):
        if "_ub_reference_" + name not in vars(cls):       # the class's own dict: idempotent, blind to inherited names
            setattr(cls, "_ub_reference_" + name, getattr(cls, name))

    keep(ActiveNerfactoModel, "get_outputs")
    ActiveNerfactoModel.get_outputs = active_nerfacto_get_outputs
    keep(NerfactoLaplaceModel, "get_outputs_unc")
    NerfactoLaplaceModel.get_outputs_unc = laplace_get_outputs_unc
    keep(NerfactoLaplaceField, "sample_laplace")
    NerfactoLaplaceField.sample_laplace = laplace_sample_laplace
    keep(NerfactoMCDropoutModel, "get_outputs_for_camera_ray_bundle")
    parent = NerfactoMCDropoutModel.__mro__[1].get_outputs_for_camera_ray_bundle     # what super() resolves to
    NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle = make_mcdropout_get_outputs(parent)
    keep(EnsemblePipeline, "get_ensemble_outputs_for_camera_ray_bundle")
    EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle = ensemble_get_outputs
    keep(ActiveSplatfactoModel, "get_outputs")
    ActiveSplatfactoModel.get_outputs = active_splatfacto_get_outputs

    if not hasattr(ref_metrics, "_ub_reference_ause"):
        ref_metrics._ub_reference_ause, ref_metrics._ub_reference_auce = ref_metrics.ause, ref_metrics.auce
    ref_metrics.ause = _ause
    ref_metrics.auce = ub_metrics.auce
    # modules that did `from nerfuncertainty.metrics import ause, auce` at import time
    for name in ("nerfuncertainty.scripts.eval_uncertainty", "nerfuncertainty.models.mcdropout.mcdropout_models",
                 "nerfuncertainty.models.ensemble.ensemble_pipeline"):
        mod = sys.modules.get(name)
        if mod is not None:
            if hasattr(mod, "ause"):
                mod.ause = _ause
            if hasattr(mod, "auce"):
                mod.auce = ub_metrics.auce
    return list(PATCHED_SURFACES)


def unpatch_reference_models() -> None:
    """Undo ``patch_reference_models`` (tests)."""
    import sys

    ref_metrics = sys.modules.get("nerfuncertainty.metrics")
    if ref_metrics is not None and hasattr(ref_metrics, "_ub_reference_ause"):
        ause0, auce0 = ref_metrics._ub_reference_ause, ref_metrics._ub_reference_auce
        for name in ("nerfuncertainty.metrics", "nerfuncertainty.scripts.eval_uncertainty",
                     "nerfuncertainty.models.mcdropout.mcdropout_models", "nerfuncertainty.models.ensemble.ensemble_pipeline"):
            mod = sys.modules.get(name)
            if mod is not None and hasattr(mod, "ause"):
                mod.ause = ause0
            if mod is not None and hasattr(mod, "auce"):
                mod.auce = auce0
        del ref_metrics._ub_reference_ause, ref_metrics._ub_reference_auce
    for modname in list(sys.modules):
        if not modname.startswith("nerfuncertainty.models"):
            continue
        for obj in vars(sys.modules[modname]).values():
            if isinstance(obj, type):
                for attr in [a for a in vars(obj) if a.startswith("_ub_reference_")]:
                    setattr(obj, attr[len("_ub_reference_"):], getattr(obj, attr))
                    delattr(obj, attr)
