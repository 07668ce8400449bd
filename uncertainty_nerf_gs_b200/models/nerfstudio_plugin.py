"""nerfstudio-side glue: put the fused CUDA path behind the reference's plugin surface.

Nothing here imports nerfstudio (or the reference package ``nerfuncertainty``) at module import time --
neither is installable in the build container -- so the rest of the package stays testable without
them.  With both installed, ``patch_reference_models()`` swaps the *eval-mode* arithmetic of the
reference's models for the ub200 kernels in place, keeping the method names, config classes, entry
points (reference pyproject.toml:18-22) and output keys untouched:

* ``ActiveNerfactoModel.get_outputs``                reference activenerfacto_model.py:83-152
* ``NerfactoLaplaceModel.get_outputs_unc``           reference laplace_model.py:456-556 (compositing part)
* ``NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle``   reference mcdropout_models.py:94-131
* ``EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle``  reference ensemble_pipeline.py:144-191
* ``ActiveSplatfactoModel`` rasterisation block      reference activesplatfacto_model.py:260-367
  (needs the tile lists; gsplat 0.1.x exposes them through ``bin_and_sort_gaussians``)
* ``nerfuncertainty.metrics.ause / auce``            reference metrics/ause.py, auce.py

Training (``self.training``) keeps the reference's autograd path: the fused compositor has no backward
yet.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import outputs as mo
from .. import metrics as ub_metrics

METHOD_NAMES = ("active-nerfacto", "active-splatfacto", "nerfacto-mcdropout", "nerfacto-laplace")
ENSEMBLE_METHOD_NAMES = ("nerfacto", "active-nerfacto", "splatfacto", "active-splatfacto")  # ensemble_utils.py:150-157


def _levels(weights_list, ray_samples_list, n):
    return [(weights_list[i], ray_samples_list[i].frustums.starts, ray_samples_list[i].frustums.ends)
            for i in range(n)]


def active_nerfacto_get_outputs(self, ray_bundle):
    """Replacement for ``ActiveNerfactoModel.get_outputs`` at eval time (activenerfacto_model.py:83-152).
    One call = one eval chunk, so chunk-wide reductions span exactly the rays of this call."""
    if self.training or self.config.predict_normals:
        return self._ub_reference_get_outputs(ray_bundle)
    from nerfstudio.field_components.field_heads import FieldHeadNames

    ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
    field_outputs = self.field.forward(ray_samples, compute_normals=False)
    fr = ray_samples.frustums
    n_rays = fr.starts.shape[0]
    flat = lambda t: t.reshape(n_rays, -1, t.shape[-1]).float()
    out = mo.active_nerfacto_outputs(
        flat(field_outputs[FieldHeadNames.DENSITY]), flat(ray_samples.deltas), flat(fr.starts), flat(fr.ends),
        flat(field_outputs[FieldHeadNames.RGB]), flat(field_outputs["rgb_var"]),
        background=_background_of(self), eval_mode=True,
        proposal_levels=[(flat(w), flat(s), flat(e)) for w, s, e in
                         _levels(weights_list, ray_samples_list, self.config.num_proposal_iterations)])
    out["density"] = field_outputs[FieldHeadNames.DENSITY]
    return out


def _background_of(model):
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
        named = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0)}
        return named[bg]
    return bg


def mcdropout_get_outputs_for_camera_ray_bundle(self, camera_ray_bundle):
    """Replacement for ``NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle``
    (mcdropout_models.py:94-131): K stochastic renders by the parent class, one fused reduce."""
    def enable_dropout(mod):
        if isinstance(mod, torch.nn.Dropout):
            mod.train()

    train_status = self.training
    if not train_status:
        self.apply(enable_dropout)
    parent = super(type(self), self).get_outputs_for_camera_ray_bundle
    outputs_list = [parent(camera_ray_bundle) for _ in range(self.config.mc_samples)]
    with torch.no_grad():
        outputs = mo.mcdropout_reduce(outputs_list)
    if not train_status:
        self.eval()
    return outputs


def ensemble_get_outputs(self, camera_ray_bundle, obb_box=None):
    """Replacement for ``EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle``
    (ensemble_pipeline.py:144-191)."""
    outputs_list = [m.get_outputs_for_camera(camera_ray_bundle, obb_box=obb_box) for m in self.models]
    tensors_only: List[Dict[str, torch.Tensor]] = [{k: v for k, v in o.items() if torch.is_tensor(v)}
                                                   for o in outputs_list]
    with torch.no_grad():
        return mo.ensemble_reduce(tensors_only)


def patch_reference_models() -> List[str]:
    """Swap the hot-path methods of the installed reference package for the ub200 ones.  Returns the list
    of patched qualified names; raises ImportError when nerfstudio / nerfuncertainty are missing."""
    import nerfuncertainty.metrics as ref_metrics
    from nerfuncertainty.models.activenerfacto.activenerfacto_model import ActiveNerfactoModel
    from nerfuncertainty.models.ensemble.ensemble_pipeline import EnsemblePipeline
    from nerfuncertainty.models.mcdropout.mcdropout_models import NerfactoMCDropoutModel

    patched = []
    ActiveNerfactoModel._ub_reference_get_outputs = ActiveNerfactoModel.get_outputs
    ActiveNerfactoModel.get_outputs = active_nerfacto_get_outputs
    patched.append("ActiveNerfactoModel.get_outputs")
    NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle = torch.no_grad()(mcdropout_get_outputs_for_camera_ray_bundle)
    patched.append("NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle")
    EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle = ensemble_get_outputs
    patched.append("EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle")

    def _ause(unc_vec, err_vec, err_type="rmse"):
        return ub_metrics.ause(unc_vec.cuda(), err_vec.cuda(), err_type)

    ref_metrics.ause = _ause
    ref_metrics.auce = ub_metrics.auce
    patched += ["nerfuncertainty.metrics.ause", "nerfuncertainty.metrics.auce"]
    # modules that did `from nerfuncertainty.metrics import ause, auce` at import time
    import sys

    for name in ("nerfuncertainty.scripts.eval_uncertainty", "nerfuncertainty.models.mcdropout.mcdropout_models",
                 "nerfuncertainty.models.ensemble.ensemble_pipeline"):
        mod = sys.modules.get(name)
        if mod is not None:
            if hasattr(mod, "ause"):
                mod.ause = _ause
            if hasattr(mod, "auce"):
                mod.auce = ub_metrics.auce
    return patched
