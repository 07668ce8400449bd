"""Load the reference's own ``ause`` / ``auce`` by file path (dev container only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` is mounted
read-only in the dev container and does NOT exist on the GPU box, so nothing that runs there
(``-m gpu`` tests, ``smoke()``, ``bench.py``) may depend on this module returning something.

Only ``nerfuncertainty/metrics/ause.py`` and ``auce.py`` are importable: everything under
``nerfuncertainty/models`` and ``nerfuncertainty/scripts`` needs nerfstudio / gsplat / backpack /
mediapy at import time, none of which is installed or installable here.  ``auce.py`` imports
``matplotlib.pyplot`` (absent) for its plotting helper only, so an empty stub module is
registered for the duration of the import.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types
from typing import Callable, Optional, Tuple

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("UB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "nerfuncertainty", "metrics", "ause.py"))


def _load(path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    assert spec is not None and spec.loader is not None
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_metrics() -> Optional[Tuple[Callable, Callable]]:
    """Return the reference's ``(ause, auce)`` callables, or ``None`` when not mounted."""
    if not reference_available():
        return None
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid  # type: ignore[attr-defined]
    stubbed = []
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            stubbed.append(name)
    try:
        base = os.path.join(REFERENCE_ROOT, "nerfuncertainty", "metrics")
        ause_mod = _load(os.path.join(base, "ause.py"), "_ub_ref_ause")
        auce_mod = _load(os.path.join(base, "auce.py"), "_ub_ref_auce")
    finally:
        for name in stubbed:
            sys.modules.pop(name, None)
    return ause_mod.ause, auce_mod.auce


@contextlib.contextmanager
def stable_torch_sort():
    """Run reference code with ``torch.sort`` forced to ``stable=True`` -- the parity contract for
    rankings (the reference's default call is not reproducible under ties on CPU)."""
    original = torch.sort

    def _stable(input, *args, **kwargs):
        kwargs["stable"] = True
        return original(input, *args, **kwargs)

    torch.sort = _stable  # type: ignore[assignment]
    try:
        yield
    finally:
        torch.sort = original  # type: ignore[assignment]
