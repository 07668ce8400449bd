"""ctypes binding of ``libub200.so`` (the C ABI declared in ``include/ub200.h``).

There is no CPU fallback: if the shared library is missing or a call fails the wrappers raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libub200.so"

UB_OK = 0
ABI_VERSION = 2
STATUS_NAMES = {0: "UB_OK", -1: "UB_ERR_BAD_ARG", -2: "UB_ERR_UNSUPPORTED", -3: "UB_ERR_WORKSPACE", -4: "UB_ERR_LAUNCH"}

UB_BG_LAST_SAMPLE, UB_BG_NONE, UB_BG_FIXED = 0, 1, 2
UB_BETA_RAW, UB_BETA_NAN_GUARD = 0, 1
UB_SPREAD_NONE, UB_SPREAD_STD, UB_SPREAD_VAR = 0, 1, 2
UB_ACT_IDENTITY, UB_ACT_SIGMOID, UB_ACT_EXP = 0, 1, 2
UB_PROLOGUE_NSUMS = 5
UB_MAX_MEMBERS = 32
UB_MAX_REDUCE_JOBS = 16
UB_MAX_COMPOSITE_BATCH = 8
UB_TILE = 16

fp = C.c_void_p  # device pointers travel as void*


class CompositeRaysArgs(C.Structure):
    _fields_ = [
        ("density", fp), ("deltas", fp), ("starts", fp), ("ends", fp), ("rgb", fp), ("beta", fp),
        ("num_rays", C.c_int64), ("num_samples", C.c_int32), ("background_mode", C.c_int32),
        ("background_rgb", C.c_float * 3), ("beta_mode", C.c_int32), ("rays_per_chunk", C.c_int64),
        ("eval_mode", C.c_int32),
        ("out_rgb", fp), ("out_accumulation", fp), ("out_depth", fp), ("out_expected_depth", fp),
        ("out_rgb_var", fp), ("out_rgb_std", fp), ("out_depth_var", fp), ("out_depth_std", fp),
        ("out_weights", fp),
    ]


class CompositeRaysBwdArgs(C.Structure):
    _fields_ = [
        ("density", fp), ("deltas", fp), ("starts", fp), ("ends", fp), ("rgb", fp), ("beta", fp),
        ("num_rays", C.c_int64), ("num_samples", C.c_int32), ("background_mode", C.c_int32),
        ("background_rgb", C.c_float * 3), ("rays_per_chunk", C.c_int64),
        ("depth", fp), ("chunk_workspace", fp),
        ("g_rgb", fp), ("g_accumulation", fp), ("g_expected_depth", fp), ("g_rgb_var", fp), ("g_rgb_std", fp),
        ("g_depth_var", fp), ("g_depth_std", fp), ("g_weights", fp),
        ("out_g_density", fp), ("out_g_rgb", fp), ("out_g_beta", fp),
    ]


class RenderWeightsArgs(C.Structure):
    _fields_ = [
        ("weights", fp), ("starts", fp), ("ends", fp),
        ("num_rays", C.c_int64), ("num_samples", C.c_int32), ("rays_per_chunk", C.c_int64),
        ("out_accumulation", fp), ("out_depth", fp), ("out_expected_depth", fp),
        ("out_depth_var", fp), ("out_depth_std", fp),
    ]


class ReduceJob(C.Structure):
    _fields_ = [
        ("members_host", C.POINTER(fp)), ("num_pixels", C.c_int64), ("channels", C.c_int32),
        ("spread_mode", C.c_int32), ("out_mean", fp), ("out_spread", fp),
    ]


class ScorePrologueArgs(C.Structure):
    _fields_ = [
        ("pred", fp), ("target", fp), ("std", fp),
        ("channels", C.c_int32), ("num_segments", C.c_int32), ("seg_offsets", fp), ("max_segment_len", C.c_int64),
        ("nll_min_std", C.c_float), ("sigma_from_var", C.c_int32),
        ("z_values", fp), ("num_z", C.c_int32),
        ("out_sq_err", fp), ("out_abs_err", fp), ("out_var", fp), ("out_sums", fp), ("out_hist", fp),
        ("out_coarse_hist", fp),
    ]


# symbol -> (restype, argtypes); also the export list the CPU test checks against include/ub200.h
SIGNATURES = {
    "ub_abi_version": (C.c_int, []),
    "ub_last_error": (C.c_char_p, []),
    "ub_sm_count": (C.c_int, []),
    "ub_composite_rays_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "ub_composite_rays": (C.c_int, [C.POINTER(CompositeRaysArgs), fp, C.c_size_t, fp]),
    "ub_composite_rays_backward": (C.c_int, [C.POINTER(CompositeRaysBwdArgs), fp]),
    "ub_composite_rays_batch_workspace_bytes": (C.c_size_t, [fp, C.c_int32]),
    "ub_composite_rays_batch": (C.c_int, [fp, C.c_int32, fp, C.c_size_t, fp]),
    "ub_render_weights_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "ub_render_weights": (C.c_int, [C.POINTER(RenderWeightsArgs), fp, C.c_size_t, fp]),
    "ub_average_sampled_weights": (C.c_int, [fp, fp, fp, fp, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, fp, fp]),
    "ub_reduce_members": (C.c_int, [C.POINTER(fp), C.c_int32, C.c_int64, C.c_int32, C.c_int32, fp, fp, fp]),
    "ub_reduce_members_batched": (C.c_int, [C.POINTER(ReduceJob), C.c_int32, C.c_int32, fp]),
    "ub_score_prologue_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64, C.c_int32]),
    "ub_score_prologue": (C.c_int, [C.POINTER(ScorePrologueArgs), fp, C.c_size_t, fp]),
    "ub_segmented_sort_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64, C.c_int64, C.c_int32]),
    "ub_segmented_sort": (C.c_int, [fp, C.c_int32, fp, C.c_int64, C.c_int64, fp, fp, fp, C.c_size_t, fp]),
    "ub_cut_prefix_sums_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64, C.c_int32, C.c_int32]),
    "ub_cut_prefix_sums": (C.c_int, [C.POINTER(fp), C.POINTER(fp), C.c_int32, C.c_int32, fp, C.c_int64,
                                     fp, C.c_int32, fp, fp, C.c_size_t, fp]),
    "ub_cut_select_sums_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32]),
    "ub_cut_select_sums": (C.c_int, [C.POINTER(fp), C.POINTER(fp), C.POINTER(fp), C.c_int32, C.c_int32, fp,
                                     C.c_int64, C.c_int64, fp, C.c_int32, fp, fp, C.c_size_t, fp]),
    "ub_cut_select_sums_ex": (C.c_int, [C.POINTER(fp), C.POINTER(fp), C.POINTER(fp), C.c_int32, C.c_int32, fp,
                                        C.c_int64, C.c_int64, fp, C.c_int32, fp, fp, fp, C.c_size_t, fp]),
    "ub_score_tail_host": (C.c_int, [fp, C.c_int32, fp, C.c_int32, fp, C.c_int32, fp, fp, C.c_int32, fp, fp,
                                     fp, fp, fp, fp, fp, fp, fp, fp]),
    "ub_laplace_ll_moments": (C.c_int, [fp, C.c_int64, C.c_int32, C.c_int32, fp, C.c_int32, C.c_int32,
                                        fp, fp, fp, fp]),
    "ub_depth_prepare_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64]),
    "ub_depth_prepare": (C.c_int, [fp, fp, fp, fp, C.c_int32, C.c_int64, fp, fp, fp, fp, fp, C.c_size_t, fp]),
    "ub_project_gaussians": (C.c_int, [fp, fp, C.c_float, fp, fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                       C.c_int32, C.c_float, C.c_int64, fp, fp, fp, fp, fp, fp, fp, fp]),
    "ub_spherical_harmonics": (C.c_int, [C.c_int32, C.c_int32, fp, fp, C.c_int64, fp, fp]),
    "ub_bin_count_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "ub_bin_count": (C.c_int, [fp, fp, C.c_int64, C.c_int32, C.c_int32, fp, fp, fp, C.c_size_t, fp]),
    "ub_bin_gaussians_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "ub_bin_gaussians": (C.c_int, [fp, fp, fp, C.c_int64, C.c_int32, C.c_int32, fp, C.c_int64, fp, fp, fp,
                                   C.c_size_t, fp]),
    "ub_composite_tiles_planes": (C.c_int, [fp, fp, fp, C.POINTER(fp), C.POINTER(C.c_int32), C.c_int32, fp, fp,
                                            C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(fp), fp, fp, fp]),
    "ub_composite_tiles_planes_backward": (C.c_int, [fp, fp, fp, C.POINTER(fp), C.POINTER(C.c_int32), C.c_int32, fp,
                                                     fp, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(fp),
                                                     fp, C.c_int64, fp, fp, fp, C.POINTER(fp), fp]),
    "ub_tile_alpha_probe": (C.c_int, [fp, fp, fp, fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, fp, fp, fp]),
    "ub_splat_normalize": (C.c_int, [fp, C.c_int32, fp, C.c_int64, C.c_int32, C.c_int32, fp, fp, fp, fp]),
    "ub_splat_depth_residual": (C.c_int, [fp, fp, fp, C.c_int32, C.c_int32, C.c_int64, fp, fp]),
    "ub_composite_tiles": (C.c_int, [fp, fp, fp, fp, C.c_int32, fp, fp, C.c_int32, C.c_int32,
                                     C.POINTER(C.c_float), fp, fp, fp]),
}


class UBError(RuntimeError):
    """A ub200 call returned a negative status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load ``libub200.so`` (built by ``uncertainty_nerf_gs_b200.build``).  Raises if it is absent --
    the product path never silently degrades to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m uncertainty_nerf_gs_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.ub_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libub200.so ABI version {got} != {ABI_VERSION} expected by this package; rebuild")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != UB_OK:
        msg = load().ub_last_error()
        raise UBError(status, msg.decode("utf8", "replace") if msg else "")
