"""nerfstudio ``method_configs`` entry points of the drop-in (``pyproject.toml`` of this repository registers them
under the reference's four entry-point names, reference pyproject.toml:18-22).

Each attribute is resolved lazily (module ``__getattr__``): the first access imports the reference's own
``MethodSpecification`` -- so method names, trainer / optimiser / model configs and descriptions are the
reference's, untouched (activenerfacto_config.py:24-61, activesplatfacto_config.py:32-90, mcdropout_configs.py:17-54,
laplace_config.py:21-58) -- after swapping the hot-path methods of the model classes those specifications point at
(``nerfstudio_plugin.patch_reference_models``).  ``ns-train active-nerfacto`` / ``ns-eval-unc`` then run the
reference's pipeline with the fused kernels underneath.  Nothing is imported from nerfstudio or the reference until
an attribute is asked for, so this module imports everywhere.
"""
from __future__ import annotations

import importlib

_SPECS = {
    # attribute here            (reference module,                                                  attribute,               method_name)
    "NerfactoMCDropoutMethod": ("nerfuncertainty.models.mcdropout.mcdropout_configs", "NerfactoMCDropoutMethod", "nerfacto-mcdropout"),
    "NerfactoLaplaceMethod": ("nerfuncertainty.models.laplace.laplace_config", "NerfactoLaplaceMethod", "nerfacto-laplace"),
    "ActiveNerfactoMethod": ("nerfuncertainty.models.activenerfacto.activenerfacto_config", "ActiveNerfactoMethod", "active-nerfacto"),
    "ActiveSplatfactoMethod": ("nerfuncertainty.models.activesplatfacto.activesplatfacto_config", "ActiveSplatfactoMethod", "active-splatfacto"),
}
ENTRY_POINTS = {"dropout": "NerfactoMCDropoutMethod", "laplace_d": "NerfactoLaplaceMethod",
                "activenerfacto": "ActiveNerfactoMethod", "activesplatfacto": "ActiveSplatfactoMethod"}
METHOD_NAMES = {attr: spec[2] for attr, spec in _SPECS.items()}


def __getattr__(name):
    spec = _SPECS.get(name)
    if spec is None:
        raise AttributeError(name)
    from .nerfstudio_plugin import patch_reference_models

    patch_reference_models()
    return getattr(importlib.import_module(spec[0]), spec[1])
