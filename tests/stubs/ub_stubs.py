"""Stand-ins for the reference's un-installable dependencies (nerfstudio 1.1.0, gsplat 0.1.11, backpack,
mediapy, matplotlib, ...), TEST INFRASTRUCTURE ONLY.

Two layers:

* ``tests/stubs/site/`` holds small *explicit* packages ``nerfstudio`` and ``gsplat`` that restate the few
  third-party pieces the hot path actually executes (``RaySamples.get_weights``, the renderers, the eval chunk
  loop, ``FieldHeadNames``, ``Cameras``, gsplat's rasterise / project entry points on top of ``oracle.splat``).
* a *permissive* meta-path finder (appended last, so real modules and the explicit stubs win) fabricates every
  other submodule of those roots on demand; attributes of such a module are dummy classes that can be
  subclassed, instantiated with any arguments, called and used as decorators.

With both installed, the reference's own modules (``nerfuncertainty.models.*``, ``scripts.eval_uncertainty``)
import unmodified in the dev container and their methods can be *executed* on CPU tensors
(``oracle/ref_exec.py``), and the plugin glue (``models/nerfstudio_plugin.py``) can run on the GPU box, where
neither nerfstudio nor the reference exists.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

SITE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "site")
STUB_ROOTS = ("nerfstudio", "gsplat", "backpack", "mediapy", "icecream", "tinycudann", "matplotlib", "mpl_toolkits",
              "tyro", "torchtyping", "jaxtyping", "torchmetrics", "cv2", "PIL", "imageio", "lpips", "torchvision",
              "pytorch_msssim", "viser", "open3d", "plotly")


if SITE not in sys.path:
    sys.path.insert(0, SITE)
from _ub_dummy import Dummy, _make  # noqa: E402,F401


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = _make(name)
        setattr(self, name, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_FINDER = None


def install() -> None:
    """Idempotent: put the explicit stubs on ``sys.path`` and append the permissive finder."""
    global _FINDER
    if SITE not in sys.path:
        sys.path.insert(0, SITE)
    if _FINDER is None:
        _FINDER = _Finder()
        sys.meta_path.append(_FINDER)


def installed() -> bool:
    return _FINDER is not None
