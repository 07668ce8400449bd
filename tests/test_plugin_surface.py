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
