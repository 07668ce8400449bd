"""Minimal stand-in for nerfstudio 1.1.0 -- TEST INFRASTRUCTURE ONLY (see ``tests/stubs/ub_stubs.py``).

Only the pieces the reference's hot path executes are restated here (from nerfstudio 1.1.0's published
behaviour; the package itself is not installable offline).  Everything else under ``nerfstudio.*`` is
fabricated on demand by the permissive finder.
"""
__version__ = "1.1.0+ub-stub"
