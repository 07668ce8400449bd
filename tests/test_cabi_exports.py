"""CPU: the C-ABI shared library builds for sm_100a, loads, exports every symbol ``include/ub200.h``
declares, and validates arguments without touching a device (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ub200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ub_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("ub_composite_rays", "ub_render_weights", "ub_reduce_members", "ub_score_prologue",
              "ub_segmented_sort", "ub_cut_prefix_sums", "ub_cut_select_sums", "ub_laplace_ll_moments", "ub_composite_tiles",
              "ub_abi_version", "ub_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_library):
    from uncertainty_nerf_gs_b200 import _lib

    lib = C.CDLL(str(built_library))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in ub200.h but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == declared_symbols(), "ctypes table out of sync with the header"
    assert _lib.load().ub_abi_version() == 2


def test_struct_layouts_match_the_header(built_library, tmp_path):
    """sizeof() of the ctypes mirrors == sizeof() of the C structs (compiled with gcc from the header)."""
    from uncertainty_nerf_gs_b200 import _lib

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ub200.h"\nint main(){printf("%zu %zu %zu\\n",'
                   "sizeof(ub_composite_rays_args),sizeof(ub_render_weights_args),sizeof(ub_score_prologue_args));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.CompositeRaysArgs), C.sizeof(_lib.RenderWeightsArgs),
                     C.sizeof(_lib.ScorePrologueArgs)]


def test_argument_validation_returns_codes_without_a_device(built_library):
    from uncertainty_nerf_gs_b200 import _lib

    lib = _lib.load()
    assert lib.ub_composite_rays(None, None, 0, None) == -1
    assert b"args is NULL" in lib.ub_last_error()
    args = _lib.CompositeRaysArgs()
    args.num_rays, args.num_samples = 8, 0
    assert lib.ub_composite_rays(C.byref(args), None, 0, None) == -1
    args.num_samples = 48
    assert lib.ub_composite_rays(C.byref(args), None, 0, None) == -1          # NULL inputs
    assert lib.ub_reduce_members(None, 5, 10, 3, 1, None, None, None) == -1
    assert lib.ub_laplace_ll_moments(None, 10, 32, 3, None, 100, 1, None, None, None, None) == -2
    assert b"hidden must be 64" in lib.ub_last_error()
    assert lib.ub_segmented_sort(None, 1, None, 0, 0, None, None, None, 0, None) == -1  # no segment table
    # chunk table (32 chunks x 16 B) + candidate flags (8 B per 8-ray tile)
    assert lib.ub_composite_rays_workspace_bytes(1 << 20, 1 << 15) == 32 * 16 + 8 * (1 << 17)
    assert lib.ub_render_weights_workspace_bytes(1 << 20, 1 << 15) == 32 * 16
    with pytest.raises(_lib.UBError):
        _lib.check(-3)


def test_only_sm100a_code_is_embedded(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", str(built_library)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs
