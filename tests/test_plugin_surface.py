"""CPU: the plugin glue imports without nerfstudio, names the reference's methods, and refuses to patch when
the reference package is not installed."""
import pytest


def test_plugin_module_imports_without_nerfstudio():
    from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as plug

    assert plug.METHOD_NAMES == ("active-nerfacto", "active-splatfacto", "nerfacto-mcdropout", "nerfacto-laplace")
    assert "active-nerfacto" in plug.ENSEMBLE_METHOD_NAMES and "nerfacto" in plug.ENSEMBLE_METHOD_NAMES
    # nerfuncertainty / nerfstudio are not installed here: a fresh interpreter must refuse to patch
    import subprocess
    import sys

    code = ("from uncertainty_nerf_gs_b200.models import nerfstudio_plugin as p\n"
            "try:\n    p.patch_reference_models()\nexcept ImportError:\n    print('refused')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    assert "refused" in out.stdout, out.stderr


def test_output_key_order_constants():
    from uncertainty_nerf_gs_b200 import pipeline
    from uncertainty_nerf_gs_b200.models import outputs

    assert outputs.STD_KEYS == ("rgb", "depth", "expected_depth")
    assert pipeline.RECORD_LEN == 2 * (6 * 100 + 5 * 99) + 25 + 3 + 1
    assert pipeline.ALL_SCALAR_KEYS[:4] == ("psnr", "ssim", "lpips", "depth_ause_mse") and pipeline.ALL_SCALAR_KEYS[-1] == "fps"
    assert pipeline.SCALAR_KEYS[:3] == ("rgb_ause_mse", "rgb_ause_mae", "rgb_ause_rmse")


def test_patch_installs_on_every_surface_of_the_real_reference():
    """Dev container only (``/root/reference`` mounted): import the reference over the stand-in nerfstudio, patch, and
    check that every class attribute the docstring promises is the ub200 replacement, the originals stay reachable,
    the entry-point module resolves to the reference's own MethodSpecifications, and unpatching restores everything."""
    from oracle import ref_exec as rx

    if not rx.available():
        pytest.skip("/root/reference not mounted")
    rx.setup()
    from uncertainty_nerf_gs_b200.models import method_configs, nerfstudio_plugin as plug

    laplace_model = rx.ref_module("models.laplace.laplace_model")
    laplace_field = rx.ref_module("models.laplace.laplace_field")
    active = rx.ref_module("models.activenerfacto.activenerfacto_model")
    splat = rx.ref_module("models.activesplatfacto.activesplatfacto_model")
    mcd = rx.ref_module("models.mcdropout.mcdropout_models")
    ens = rx.ref_module("models.ensemble.ensemble_pipeline")
    ev = rx.ref_module("scripts.eval_uncertainty")
    import nerfuncertainty.metrics as ref_metrics

    originals = {"active": active.ActiveNerfactoModel.get_outputs, "unc": laplace_model.NerfactoLaplaceModel.get_outputs_unc,
                 "sample": laplace_field.NerfactoLaplaceField.sample_laplace, "splat": splat.ActiveSplatfactoModel.get_outputs,
                 "mcd": mcd.NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle,
                 "ens": ens.EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle, "ause": ref_metrics.ause}
    try:
        patched = plug.patch_reference_models()
        assert patched == list(plug.PATCHED_SURFACES) and len(patched) == 8
        assert active.ActiveNerfactoModel.get_outputs is plug.active_nerfacto_get_outputs
        assert laplace_model.NerfactoLaplaceModel.get_outputs_unc is plug.laplace_get_outputs_unc
        assert laplace_field.NerfactoLaplaceField.sample_laplace is plug.laplace_sample_laplace
        assert splat.ActiveSplatfactoModel.get_outputs is plug.active_splatfacto_get_outputs
        assert ens.EnsemblePipeline.get_ensemble_outputs_for_camera_ray_bundle is plug.ensemble_get_outputs
        assert mcd.NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle is not originals["mcd"]
        assert active.ActiveNerfactoModel._ub_reference_get_outputs is originals["active"]
        assert laplace_field.NerfactoLaplaceField._ub_reference_sample_laplace is originals["sample"]
        # forward_unc / get_outputs of the field stay the reference's: the default-argument quirk (:516-520) is theirs
        assert "forward_unc" in vars(laplace_field.NerfactoLaplaceField) and not hasattr(plug, "laplace_forward_unc")
        # modules that imported ause / auce by name see the replacements
        assert ev.ause is plug._ause and ref_metrics.auce.__module__.endswith("uncertainty_nerf_gs_b200.metrics")
        # entry points: lazily resolved, the reference's own specification objects
        for attr, method_name in method_configs.METHOD_NAMES.items():
            spec = getattr(method_configs, attr)
            ref_spec = getattr(__import__(method_configs._SPECS[attr][0], fromlist=["x"]), attr)
            assert spec is ref_spec
        assert sorted(method_configs.METHOD_NAMES.values()) == sorted(plug.METHOD_NAMES)
        plug.patch_reference_models()                                   # idempotent: originals are not overwritten
        assert active.ActiveNerfactoModel._ub_reference_get_outputs is originals["active"]
    finally:
        plug.unpatch_reference_models()
    assert active.ActiveNerfactoModel.get_outputs is originals["active"]
    assert laplace_model.NerfactoLaplaceModel.get_outputs_unc is originals["unc"]
    assert splat.ActiveSplatfactoModel.get_outputs is originals["splat"]
    assert mcd.NerfactoMCDropoutModel.get_outputs_for_camera_ray_bundle is originals["mcd"]
    assert ref_metrics.ause is originals["ause"] and ev.ause is originals["ause"]
    assert not hasattr(active.ActiveNerfactoModel, "_ub_reference_get_outputs")


def test_pyproject_registers_the_reference_entry_point_names():
    import pathlib
    import re

    root = pathlib.Path(__file__).resolve().parents[1]
    text = (root / "pyproject.toml").read_text()
    ours = dict(re.findall(r"^(\w+) = '(uncertainty_nerf_gs_b200[^']+)'", text, flags=re.M))
    from uncertainty_nerf_gs_b200.models import method_configs

    assert set(ours) == set(method_configs.ENTRY_POINTS)                # dropout, laplace_d, activenerfacto, activesplatfacto
    for name, target in ours.items():
        mod, attr = target.split(":")
        assert mod == "uncertainty_nerf_gs_b200.models.method_configs" and attr == method_configs.ENTRY_POINTS[name]
    ref = pathlib.Path("/root/reference/pyproject.toml")
    if ref.exists():
        theirs = dict(re.findall(r"^(\w+) = 'nerfuncertainty\.models[^:]+:(\w+)'", ref.read_text(), flags=re.M))
        assert theirs == {k: v for k, v in method_configs.ENTRY_POINTS.items()}
