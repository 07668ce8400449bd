"""pytest configuration: the ``gpu`` marker and import paths.

``-m "not gpu"`` runs everywhere (oracle vs golden vectors, host logic, C-ABI exports, gloo);
``-m gpu`` needs a B200 and is the parity suite proper (CUDA path vs oracle through the C ABI).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_library():
    """Build (or reuse) libub200.so; every test that touches the C ABI depends on it."""
    from uncertainty_nerf_gs_b200.build import build_library

    return build_library()
