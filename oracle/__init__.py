"""CPU oracle for the uncertainty rendering-and-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it is a
CPU (torch-CPU / numpy / scipy) restatement of the arithmetic the reference
(AaltoML/uncertainty-nerf-gs, ``/root/reference``) performs on this path, and
exists solely to *check* the CUDA implementation.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product package
(``uncertainty_nerf_gs_b200``) never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):

* ``oracle.metrics.ause`` / ``oracle.metrics.auce`` -- PINNED: checked against the
  reference's own ``nerfuncertainty/metrics/{ause,auce}.py`` executed in the dev
  container (``tests/golden/make_golden.py`` imports them by file path) and
  against the committed golden vectors in ``tests/golden/``.
* ``oracle.compositing`` / ``oracle.reduce`` / ``oracle.laplace`` / ``oracle.splat``
  -- PARITY UNPINNED at the third-party boundary: the reference delegates this
  arithmetic to nerfstudio 1.1.0 / gsplat 0.1.11, which are neither vendored in
  ``/root/reference`` nor installable here, and the reference has no tests or
  golden vectors.  These modules restate the published algorithms of those
  dependencies and are anchored on the reference's own call sites (cited per
  function) and on the in-repo restatement ``laplace_model.py:47-62,102-107``.
"""
