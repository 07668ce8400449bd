"""Minimal stand-in for gsplat 0.1.11 -- TEST INFRASTRUCTURE ONLY (see ``tests/stubs/ub_stubs.py``): the three entry
points the reference calls (activesplatfacto_model.py:12-15, 221-234, 245, 260-355), implemented by the CPU oracle
``oracle.splat`` (published gsplat 0.1.11 algorithms).  Keeps gsplat's call signatures, including the internal
re-binning of every ``rasterize_gaussians`` call."""
__version__ = "0.1.11+ub-stub"
