"""CPU oracle for the uncertainty rendering-and-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it is a
CPU (torch-CPU / numpy / scipy) restatement of the arithmetic the reference
(AaltoML/uncertainty-nerf-gs, ``/root/reference``) performs on this path, and
exists solely to *check* the CUDA implementation.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product package
(``uncertainty_nerf_gs_b200``) never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"; the tests are ``tests/test_oracle_pinned.py`` and
``tests/test_oracle_golden.py``):

* Every restatement of the REFERENCE'S OWN code is PINNED: ``oracle/ref_exec.py`` imports the reference's
  modules unmodified (over the stand-in third-party packages of ``tests/stubs/site``) and executes their
  methods -- ``ComputeWeightsModule`` / ``SumModule``, ``ActiveNerfactoModel.get_outputs`` and its chunk
  loop, ``NerfactoLaplaceModel.get_outputs_unc``, ``NerfactoLaplaceField.sample_laplace``, the MC-dropout
  and ensemble reduces, ``ActiveSplatfactoModel.get_outputs``, ``get_unc_metrics_rgb / _depth``,
  ``negative_gaussian_loglikelihood``, the test-set loop with its ``.npy`` dumps, ``ause`` / ``auce`` -- and
  ``oracle.compositing / reduce / laplace / splat / metrics`` equal them BIT FOR BIT on the same inputs
  (live in the dev container, and everywhere against the committed ``tests/golden/ref_*.npz`` those
  executions produced, script ``tests/golden/make_golden.py``).
* What stays a restatement is only what lives inside the un-vendored third-party dependencies the
  reference calls into: nerfstudio 1.1.0's renderers / ``RaySamples.get_weights`` / eval chunk loop
  (``tests/stubs/site/nerfstudio``; anchored on the reference's in-repo copy ``laplace_model.py:47-62,
  102-107``, which is executed and agrees bit for bit) and gsplat 0.1.11's ``rasterize_gaussians`` /
  ``project_gaussians`` / ``spherical_harmonics`` (``oracle/splat.py`` behind ``tests/stubs/site/gsplat``;
  published algorithms, neither package is installable here).
"""
