"""Execute the REFERENCE'S OWN hot-path code on CPU tensors (dev container only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` is mounted read-only in the dev
container and absent on the GPU box, so nothing that runs there may depend on this module; it is used by
``tests/test_oracle_pinned.py`` (``-m "not gpu"``, skipped when the reference is not mounted) to pin the oracle
restatements bit-for-bit, and by ``tests/golden/make_golden.py`` to generate the committed golden vectors.

How: the reference's modules need nerfstudio / gsplat / backpack / mediapy / matplotlib at import time, none of
which is installable here.  ``tests/stubs/ub_stubs.install()`` provides stand-ins (small explicit packages for
the third-party pieces the path executes, a permissive finder for the rest), after which the reference modules
import UNMODIFIED and their methods run as written:

* ``ComputeWeightsModule`` / ``SumModule``                            laplace_model.py:47-62, 102-107
* ``ActiveNerfactoModel.get_outputs`` (+ the inherited chunk loop)     activenerfacto_model.py:83-152
* ``NerfactoLaplaceModel.get_outputs_unc`` / ``..._ray_bundle_unc``    laplace_model.py:417-556
* ``NerfactoLaplaceField.sample_laplace``                              laplace_field.py:528-568
* ``NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle``         mcdropout_models.py:94-131
* ``EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle``      ensemble_pipeline.py:144-191
* ``ActiveSplatfactoModel.get_outputs``                                activesplatfacto_model.py:142-367
* ``get_unc_metrics_rgb`` / ``get_unc_metrics_depth`` / ``negative_gaussian_loglikelihood`` /
  ``get_image_metrics_and_images_unc`` / ``get_average_uncertainty_metrics``   eval_uncertainty.py:306-1079
* ``ause`` / ``auce`` / ``plot_auce_curves``                           metrics/ause.py, auce.py

What remains restated rather than executed is only what lives in the un-vendored dependencies themselves:
nerfstudio's renderers / ``RaySamples.get_weights`` / chunk loop (``tests/stubs/site/nerfstudio``) and gsplat's
rasteriser (``oracle/splat.py`` behind ``tests/stubs/site/gsplat``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ref_loader

Tensor = torch.Tensor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")

_READY = False


def available() -> bool:
    return ref_loader.reference_available() and os.path.isdir(STUBS)


def setup() -> None:
    """Install the stand-ins and put the reference on ``sys.path`` (idempotent)."""
    global _READY
    if _READY:
        return
    if not available():
        raise RuntimeError("/root/reference is not mounted: the reference cannot be executed here")
    if STUBS not in sys.path:
        sys.path.insert(0, STUBS)
    import ub_stubs

    ub_stubs.install()
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid  # type: ignore[attr-defined]
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.append(ref_loader.REFERENCE_ROOT)
    _READY = True


def ref_module(name: str):
    """Import ``nerfuncertainty.<name>`` from the mounted reference."""
    setup()
    return importlib.import_module("nerfuncertainty." + name)


def fakes():
    setup()
    import fakes as f

    return f


# --------------------------------------------------------------------------------------------------------
# compositing
def compute_weights(density: Tensor, deltas: Tensor) -> Tensor:
    return ref_module("models.laplace.laplace_model").ComputeWeightsModule()(density, deltas)


def sum_module(x: Tensor, w: Tensor) -> Tensor:
    return ref_module("models.laplace.laplace_model").SumModule()(x, w)


def make_active_nerfacto(inp: Dict[str, Tensor], proposal_levels=(), background="last_sample", training=False,
                         eval_num_rays_per_chunk: int = 1 << 15):
    mod = ref_module("models.activenerfacto.activenerfacto_model")
    cfg = mod.ActiveNerfactoModelConfig(background_color=background, eval_num_rays_per_chunk=eval_num_rays_per_chunk,
                                        num_proposal_iterations=len(proposal_levels))
    # populate_modules of the reference builds the hash-grid field (a producer, out of scope): construct the
    # object without it and attach the stand-in producers instead
    model = mod.ActiveNerfactoModel.__new__(mod.ActiveNerfactoModel)
    _init_nerfacto_family(model, cfg)
    from nerfstudio.model_components.renderers import UncertaintyRenderer

    model.renderer_uncertainty = UncertaintyRenderer()
    fakes().attach_producers(model, inp, proposal_levels)
    model.train(training)
    return model


def _init_nerfacto_family(model, cfg):
    from nerfstudio.models.base_model import Model
    from nerfstudio.models.nerfacto import NerfactoModel

    torch.nn.Module.__init__(model)
    model.config = cfg
    model.collider = None
    model.kwargs = {}
    model.device_indicator_param = torch.nn.Parameter(torch.empty(0))
    NerfactoModel.populate_modules(model)
    assert isinstance(model, Model)


def active_nerfacto_get_outputs(inp: Dict[str, Tensor], proposal_levels=(), background="last_sample",
                                training: bool = False) -> Dict[str, Tensor]:
    """``ActiveNerfactoModel.get_outputs`` on one ray batch (= one eval chunk)."""
    model = make_active_nerfacto(inp, proposal_levels, background, training)
    bundle = fakes().flat_ray_bundle(inp["density"].shape[0])
    with torch.set_grad_enabled(training):
        return model.get_outputs(bundle)


def active_nerfacto_camera(inp: Dict[str, Tensor], height: int, width: int, proposal_levels=(),
                           background="last_sample", chunk: int = 1 << 15) -> Dict[str, Tensor]:
    """``get_outputs_for_camera_ray_bundle`` (the inherited chunk loop) of the active-nerfacto model."""
    model = make_active_nerfacto(inp, proposal_levels, background, False, chunk)
    return model.get_outputs_for_camera_ray_bundle(fakes().camera_ray_bundle(height, width))


def make_laplace(inp: Dict[str, Tensor], density_var: Optional[Tensor], proposal_levels=(),
                 background="last_sample", chunk: int = 1 << 15):
    mod = ref_module("models.laplace.laplace_model")
    cfg = mod.NerfactoLaplaceModelConfig(background_color=background, eval_num_rays_per_chunk=chunk,
                                         num_proposal_iterations=len(proposal_levels))
    model = mod.NerfactoLaplaceModel.__new__(mod.NerfactoLaplaceModel)
    _init_nerfacto_family(model, cfg)
    from nerfstudio.model_components.renderers import UncertaintyRenderer

    model.uncertainty_renderer = UncertaintyRenderer()
    model.rgb_la_renderer = mod.SumModule()
    f = fakes()
    field = f.TensorField(inp["density"], inp["rgb"], rgb_var=inp["beta"], density_var=density_var)
    f.attach_producers(model, inp, proposal_levels, field=field)
    model.eval()
    return model


def laplace_get_outputs_unc(inp: Dict[str, Tensor], density_var: Optional[Tensor] = None,
                            use_deterministic_density: bool = True, proposal_levels=(), background="last_sample",
                            seed: Optional[int] = None) -> Dict[str, Tensor]:
    """``NerfactoLaplaceModel.get_outputs_unc`` on one ray batch.  ``inp['rgb']`` / ``inp['beta']`` play the
    last-layer Laplace moments ``mu_rgb`` / ``rgb_var`` the field returns.  With sampled density the 100 draws come
    from torch's global generator exactly as in the reference; ``seed`` seeds it."""
    model = make_laplace(inp, density_var, proposal_levels, background)
    bundle = fakes().flat_ray_bundle(inp["density"].shape[0])
    if seed is not None:
        torch.manual_seed(seed)
    with torch.no_grad():
        return model.get_outputs_unc(bundle, is_inference=True, use_deterministic_density=use_deterministic_density)


def sample_laplace(linear: torch.nn.Linear, activation, diag_ggn: Tensor, x: Tensor, n_samples: int = 100,
                   prior_prec: float = 1.0, eps: float = 1e-9, seed: int = 0):
    """``NerfactoLaplaceField.sample_laplace`` (it never touches ``self``).  Returns ``(mu, sigma2)``; the draws
    are ``torch.randn(n_samples, n_params)`` on the global generator seeded with ``seed``."""
    mod = ref_module("models.laplace.laplace_field")
    torch.manual_seed(seed)
    return mod.NerfactoLaplaceField.sample_laplace(None, module=linear, activation=activation, diag_ggn=diag_ggn,
                                                   input=x, n_samples=n_samples, prior_prec=prior_prec, eps=eps)


# --------------------------------------------------------------------------------------------------------
# across-pass reduces
def mcdropout_reduce(outputs_list: Sequence[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    """``NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle`` with the K stochastic parent renders replayed
    from ``outputs_list`` (the zero-argument ``super()`` call resolves to the stand-in nerfacto parent, whose
    ``get_outputs_for_camera_ray_bundle`` is replaced for the duration of the call)."""
    mod = ref_module("models.mcdropout.mcdropout_models")
    from nerfstudio.models.nerfacto import NerfactoModel

    cfg = mod.NerfactoMCDropoutModelConfig(mc_samples=len(outputs_list))
    model = mod.NerfactoMCDropoutModel.__new__(mod.NerfactoMCDropoutModel)
    _init_nerfacto_family(model, cfg)
    model.eval()
    replay = fakes().ReplayModel(outputs_list)
    saved = NerfactoModel.get_outputs_for_camera_ray_bundle
    NerfactoModel.get_outputs_for_camera_ray_bundle = lambda self, b: replay.get_outputs_for_camera_ray_bundle(b)
    try:
        return model.get_outputs_for_camera_ray_bundle(None)
    finally:
        NerfactoModel.get_outputs_for_camera_ray_bundle = saved


def ensemble_reduce(outputs_list: Sequence[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    """``EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle`` with M replayed member models."""
    mod = ref_module("models.ensemble.ensemble_pipeline")
    f = fakes()
    self = types.SimpleNamespace(models=[f.ReplayModel([o]) for o in outputs_list])
    return mod.EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle(self, None, obb_box=None)


# --------------------------------------------------------------------------------------------------------
# splat
def make_active_splatfacto(gauss: Dict[str, Tensor], log_unc: Tensor, background=(0.1, 0.2, 0.3), sh_degree: int = 3):
    mod = ref_module("models.activesplatfacto.activesplatfacto_model")
    cfg = mod.ActiveSplatfactoModelConfig(sh_degree=sh_degree, background_color="random")
    model = mod.ActiveSplatfactoModel(cfg, seed_gaussians=gauss)
    model.gauss_params["log_uncertainties"] = torch.nn.Parameter(log_unc.clone())
    model.background_color = torch.tensor(background)
    model.eval()
    return model


def active_splatfacto_get_outputs(gauss: Dict[str, Tensor], log_unc: Tensor, camera, background=(0.1, 0.2, 0.3),
                                  sh_degree: int = 3) -> Dict[str, Tensor]:
    """``ActiveSplatfactoModel.get_outputs(camera)`` in eval mode: projection, SH colours and the four
    rasterisation passes go through the gsplat stand-in (= ``oracle.splat``), everything else is the reference."""
    model = make_active_splatfacto(gauss, log_unc, background, sh_degree)
    with torch.no_grad():
        out = model.get_outputs(camera)
    out["_xys"], out["_radii"] = model.xys, model.radii
    return out


# --------------------------------------------------------------------------------------------------------
# scoring
def nll(preds: Tensor, targets: Tensor, stds: Tensor, eps: float) -> Tensor:
    return ref_module("scripts.eval_uncertainty").negative_gaussian_loglikelihood(preds, targets, stds, eps=eps)


class _ScoreModel:
    """The few members of ``pipeline.model`` the scoring functions read (eval_uncertainty.py:320-322, 659-700)."""

    def __init__(self, out_dir: Path, dataset_path: Path, image_metrics=(0.0, 0.0, 0.0)):
        self.device = torch.device("cpu")
        self.output_path = Path(out_dir) / "output.json"
        self.dataset_path = dataset_path
        self._im = image_metrics

    def psnr(self, a, b):
        return torch.tensor(self._im[0])

    def ssim(self, a, b):
        return torch.tensor(self._im[1])

    def lpips(self, a, b):
        return torch.tensor(self._im[2])

    @staticmethod
    def get_gt_img(image):
        return image.float() / 255.0 if image.dtype == torch.uint8 else image

    @staticmethod
    def composite_with_background(image, background):
        if image.shape[2] == 4:
            alpha = image[..., -1].unsqueeze(-1).repeat((1, 1, 3))
            return alpha * image[..., :3] + (1 - alpha) * background
        return image


def unc_metrics_rgb(outputs: Dict[str, Tensor], gt: Tensor, min_rgb_std_for_nll: float = 3e-2,
                    stable: bool = True) -> Dict[str, object]:
    """``get_unc_metrics_rgb``; ``stable`` forces ``torch.sort(stable=True)`` inside ``ause`` (the ranking
    contract; the reference's default CPU sort is not reproducible under ties)."""
    ev = ref_module("scripts.eval_uncertainty")
    model = _ScoreModel(Path("."), Path("."))
    import contextlib

    ctx = ref_loader.stable_torch_sort() if stable else contextlib.nullcontext()
    with ctx:
        return ev.get_unc_metrics_rgb(model, 0, {"image": gt}, outputs, Path("."), Path("."),
                                      min_rgb_std_for_nll=min_rgb_std_for_nll)


def write_depth_side_inputs(dataset_path: Path, depth_gts: Sequence[np.ndarray], scale: float) -> None:
    """The files ``get_unc_metrics_depth`` loads (eval_uncertainty.py:432-436)."""
    dataset_path.mkdir(parents=True, exist_ok=True)
    np.savetxt(str(dataset_path) + "/scale_parameters.txt", np.array([scale]), delimiter=",")
    for i, g in enumerate(depth_gts):
        np.save(str(dataset_path) + "/depth_gt_{:02d}.npy".format(i), g)


def unc_metrics_depth(outputs: Dict[str, Tensor], img_num: int, dataset_path: Path, out_dir: Path,
                      min_depth_std_for_nll: float = 1.0, stable: bool = True) -> Dict[str, object]:
    ev = ref_module("scripts.eval_uncertainty")
    import contextlib

    ctx = ref_loader.stable_torch_sort() if stable else contextlib.nullcontext()
    with ctx:
        return ev.get_unc_metrics_depth(img_num, outputs, dataset_path, Path(out_dir),
                                        min_depth_std_for_nll=min_depth_std_for_nll)


def average_uncertainty_metrics(views: Sequence[Dict[str, Tensor]], gts: Sequence[Tensor], out_dir: Path,
                                dataset_path: Optional[Path] = None, eval_depth: bool = False,
                                image_metrics=(0.0, 0.0, 0.0), min_rgb_std_for_nll: float = 3e-2,
                                min_depth_std_for_nll: float = 1.0, stable: bool = True) -> Dict[str, float]:
    """``get_average_uncertainty_metrics``: the reference's per-view loop, curve accumulation, ``plot_auce_curves``
    (which writes the ``auce_{output}_*.npy`` files into ``out_dir/plots``) and the final float32 means, with the
    model renders replayed from ``views``."""
    ev = ref_module("scripts.eval_uncertainty")
    from nerfstudio.cameras.cameras import Cameras
    import contextlib

    out_dir = Path(out_dir)
    (out_dir / "plots").mkdir(parents=True, exist_ok=True)
    model = _ScoreModel(out_dir, dataset_path, image_metrics)
    loader = []
    for v, g in zip(views, gts):
        h, w = v["rgb"].shape[:2]
        loader.append((Cameras(torch.eye(4)[:3], 1.0, 1.0, w / 2, h / 2, w, h), {"image": g}))
    it = iter(views)
    self = types.SimpleNamespace(datamanager=types.SimpleNamespace(fixed_indices_eval_dataloader=loader), model=model)
    ctx = ref_loader.stable_torch_sort() if stable else contextlib.nullcontext()
    with ctx:
        return ev.get_average_uncertainty_metrics(self, lambda camera: next(it), eval_depth_unc=eval_depth,
                                                  eval_rgb_unc=True, min_rgb_std_for_nll=min_rgb_std_for_nll,
                                                  min_depth_std_for_nll=min_depth_std_for_nll)
