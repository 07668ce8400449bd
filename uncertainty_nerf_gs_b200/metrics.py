"""Drop-in ``ause`` / ``auce`` (same signatures and return values as ``nerfuncertainty.metrics``) and the
batched per-image scorer, running on the ub200 CUDA kernels.

Mirrors reference ``nerfuncertainty/metrics/ause.py:7-44``, ``auce.py:10-57`` and
``scripts/eval_uncertainty.py:306-402`` (``get_unc_metrics_rgb``).  The device does the O(N) work
(stable segmented radix sort, float64 cut-point prefix sums, metric prologue, NLL, interval
histogram); the O(100) tail (normalisation, trapezoid areas, dict assembly) stays on the host in
numpy exactly as the reference computes it, so dtypes and rounding of the returned objects match.

Ranking contract: ``torch.sort(stable=True)`` order (the reference's default unstable CPU sort is not
reproducible under ties).  AUCE contract: NumPy >= 2 promotion (float64 interval arithmetic).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.stats
import torch

from . import ops

Tensor = torch.Tensor

if not hasattr(np, "trapz"):  # newer NumPy dropped the alias the reference calls
    np.trapz = np.trapezoid  # type: ignore[attr-defined]

_Z_CACHE: Dict[str, Tensor] = {}


_RATIOS = np.linspace(0, 1, 100, endpoint=False)  # ause.py:8
_RATIOS.setflags(write=False)
N_RATIOS = len(_RATIOS)


def _ratios() -> np.ndarray:
    return _RATIOS


def ause_cut_counts(n: int) -> np.ndarray:
    """``int((1 - r) * n)`` for the 100 removal ratios, evaluated in float64 with truncation exactly as
    ause.py:16,30 does (NOT ``n * (100 - i) // 100``: the two differ at ~25 of 100 indices)."""
    c = _CUTS.get(n)
    if c is None:
        if len(_CUTS) > 256:
            _CUTS.clear()
        c = _CUTS[n] = np.array([int((1 - r) * n) for r in _ratios()], dtype=np.int64)
    return c


_ALPHAS: Optional[List[np.float64]] = None
_Z_HOST: Optional[np.ndarray] = None
_CUTS: Dict[int, np.ndarray] = {}


def _alphas() -> List[np.float64]:
    global _ALPHAS
    if _ALPHAS is None:
        _ALPHAS = list(np.arange(start=0.01, stop=1.0, step=0.01))  # auce.py:17
    return _ALPHAS


def z_values_host() -> np.ndarray:
    """``scipy.stats.norm.ppf(1 - alpha/2)`` for the 99 alphas (auce.py:21-22); computed once -- the scalar
    scipy call costs tens of microseconds and the reference pays it 198 times per image."""
    global _Z_HOST
    if _Z_HOST is None:
        _Z_HOST = np.array([scipy.stats.norm.ppf(1.0 - a / 2) for a in _alphas()], dtype=np.float64)
    return _Z_HOST


def _z_table(device) -> Tensor:
    key = str(device)
    if key not in _Z_CACHE:
        _Z_CACHE[key] = torch.from_numpy(z_values_host()).to(device)
    return _Z_CACHE[key]


def _prefix_means(sums: np.ndarray, cuts: np.ndarray, err_type: str) -> np.ndarray:
    """float32 value of ``err_sorted[:c].mean()`` (and ``torch.sqrt`` of it for rmse) from the float64
    prefix sums; an empty slice gives NaN like torch."""
    with np.errstate(divide="ignore", invalid="ignore"):
        m = np.where(cuts > 0, sums / np.maximum(cuts, 1), np.nan).astype(np.float32)
        if err_type == "rmse":
            m = np.sqrt(m)
    return m


def _py_max(arr: np.ndarray):
    """Python's ``max(iterable)`` (first maximal element; a NaN never wins a comparison but a leading NaN
    sticks), which is what ause.py:37 applies to both curves."""
    if not np.isnan(arr).any():
        return arr.max()
    best = arr[0]
    for v in arr[1:]:
        if v > best:
            best = v
    return best


_RATIO_STEPS = np.diff(_RATIOS)


def _trapz_rows(y: np.ndarray, dx: np.ndarray) -> np.ndarray:
    """``np.trapz(y, x, axis=-1)`` with ``dx = np.diff(x)`` precomputed: the same expression numpy evaluates
    (``(d * (y[1:] + y[:-1]) / 2.0).sum(-1)``, pairwise row sums), without its per-call diff / warning cost."""
    return (dx * (y[..., 1:] + y[..., :-1]) / 2.0).sum(-1)


def _py_max_rows(arr: np.ndarray) -> np.ndarray:
    """``_py_max`` of every row of ``arr [B, n]``."""
    out = arr.max(axis=1)
    bad = np.isnan(out)
    for i in np.nonzero(bad)[0]:
        out[i] = _py_max(arr[i])
    return out


def _ause_tail_batch(oracle_f32: np.ndarray, by_unc_f32: np.ndarray):
    """ause.py:27-44 downstream of the slice means for ``[B, 100]`` curves at once: normalise both curves by
    the common maximum and integrate the gap.  dtypes follow the reference row by row: the oracle curve is
    built from float32 scalars, the by-uncertainty curve lives in a float64 array, and ``max(a, b)`` keeps
    ``a`` (np.float32) unless ``b > a`` (np.float64) -- which decides whether the oracle curve is divided in
    float32 or float64.  Returns ``(oracle_curves list, by_unc_curves [B,100] f64, ause [B] f64)``."""
    by_unc = by_unc_f32.astype(np.float64)
    a = _py_max_rows(oracle_f32)                      # float32
    b = _py_max_rows(by_unc)                          # float64
    b_wins = b > a
    with np.errstate(divide="ignore", invalid="ignore"):
        max64 = np.where(b_wins, b, a.astype(np.float64))
        oracle32 = oracle_f32 / a[:, None]                                    # float32 / np.float32
        oracle64 = oracle_f32.astype(np.float64) / max64[:, None]             # float32 / np.float64 -> float64
        by_unc = by_unc / max64[:, None]
        gap = by_unc - np.where(b_wins[:, None], oracle64, oracle32.astype(np.float64))
        ause_vals = _trapz_rows(gap, _RATIO_STEPS)
    oracle = [oracle64[i] if b_wins[i] else oracle32[i] for i in range(len(a))]
    return oracle, by_unc, ause_vals


def _ause_tail(oracle_f32: np.ndarray, by_unc_f32: np.ndarray):
    """Single-curve form of ``_ause_tail_batch`` with the reference's return tuple."""
    o, b, a = _ause_tail_batch(np.asarray(oracle_f32)[None], np.asarray(by_unc_f32)[None])
    return _RATIOS, o[0], b[0], a[0]


_SELECT_MAX_LEN = 1 << 24


def _use_select(max_len: int) -> bool:
    """The AUSE prefix sums come from the sort-free multi-cut select (``ub_cut_select_sums``) unless
    ``UB_AUSE_SORT=1`` asks for the full segmented sort (same sums up to float64 summation order; the sort
    remains the path that returns the permutation, ``ops.segmented_sort``)."""
    return os.environ.get("UB_AUSE_SORT", "0") != "1" and max_len <= _SELECT_MAX_LEN


def _ause_sums(vec: Tensor, lens, cuts: np.ndarray, coarse: Optional[Tensor] = None,
               out: Optional[Tensor] = None) -> Tensor:
    """``[B, 4, ncuts]`` float64: payload sums under every cut for (abs err by var, sq err by var, abs err
    ascending, sq err ascending) from the prologue's ``[3, total]`` buffer (var, abs err, sq err); ``coarse``: the
    prologue's key histograms of those three vectors, in that order; ``out``: where to write them."""
    ae, se = vec[1], vec[2]
    if _use_select(max(lens) if len(lens) else 0):
        return ops.cut_select_sums([(vec[0], ae, se), (ae, ae, None), (se, se, None)], lens, cuts, coarse=coarse,
                                   out=out)
    total = vec.shape[1]
    sorted_all, perm_all = ops.segmented_sort(vec.reshape(-1), list(lens) * 3, want_perm=True, want_keys=True)
    perm_var = perm_all[:total]
    sums = ops.cut_prefix_sums([ae, se, sorted_all[total:2 * total], sorted_all[2 * total:]],
                               [perm_var, perm_var, None, None], lens, cuts)
    if out is not None:
        out.copy_(sums)
        return out
    return sums


def _check_err_type(err_type: str) -> None:
    if err_type not in ("rmse", "mse", "mae"):
        raise ValueError(f"err_type must be 'rmse', 'mse' or 'mae', got {err_type!r}")


def ause(unc_vec: Tensor, err_vec: Tensor, err_type: str = "rmse"
         ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.float64]:
    """Drop-in for ``nerfuncertainty.metrics.ause`` (ause.py:7-44): ``(ratio_removed, ause_err,
    ause_err_by_var, ause)``.  ``unc_vec`` / ``err_vec`` are 1-D CUDA float32 tensors."""
    _check_err_type(err_type)
    if unc_vec.dim() != 1 or unc_vec.shape != err_vec.shape:
        raise ValueError("unc_vec and err_vec must be 1-D tensors of equal length")
    n = len(err_vec)
    cuts = ause_cut_counts(n)
    if _use_select(n):
        unc32, err32 = unc_vec.reshape(-1).to(torch.float32), err_vec.reshape(-1).to(torch.float32)
        sums = ops.cut_select_sums([(err32, err32, None), (unc32, err32, None)], [n], cuts[None, :])
    else:
        unc32 = unc_vec.reshape(-1).to(torch.float32)
        err32 = err_vec.reshape(-1).to(torch.float32).contiguous()
        both = torch.stack([unc32, err32])
        sorted_all, perm_all = ops.segmented_sort(both.reshape(-1), [n, n], want_perm=True, want_keys=True)
        sums = ops.cut_prefix_sums([sorted_all[n:], err32], [None, perm_all[:n]], [n], cuts[None, :])
    host = sums[0].cpu().numpy()
    return _ause_tail(_prefix_means(host[0], cuts, err_type), _prefix_means(host[1], cuts, err_type))


_ALPHA_ARR: Optional[np.ndarray] = None
_ONE_MINUS_ALPHA: Optional[np.ndarray] = None
_ALPHA_STEPS: Optional[np.ndarray] = None


def _alpha_tables() -> None:
    global _ALPHA_ARR, _ONE_MINUS_ALPHA, _ALPHA_STEPS
    if _ALPHA_ARR is None:
        _ALPHA_ARR = np.array(_alphas())
        _ONE_MINUS_ALPHA = 1.0 - _ALPHA_ARR
        _ALPHA_STEPS = np.diff(_ALPHA_ARR)


def _auce_from_hist_batch(hist: np.ndarray, sigma_sum: np.ndarray, n: np.ndarray, z: np.ndarray) -> List[Dict[str, object]]:
    """auce.py:24-54 from the interval histograms ``hist [B, nz+1]``: coverage_k = #{elements satisfying > k
    thresholds} / n; mean interval length = 2 z_k mean(sigma) (equal to the reference's float64
    ``mean(upper - lower)`` to ~2e-16 relative).  ``sigma_sum, n``: per-image float64."""
    _alpha_tables()
    inside = np.cumsum(hist[:, ::-1], axis=1)[:, ::-1][:, 1:]  # count with c > k, k = 0..nz-1
    with np.errstate(divide="ignore", invalid="ignore"):
        coverage = inside.astype(np.float64) / n[:, None]
        avg_len = (2.0 * z)[None, :] * (sigma_sum / n)[:, None]
    err = coverage - _ONE_MINUS_ALPHA
    abs_err = np.abs(err)
    neg_err = (abs_err - err) / 2.0
    auc_len = _trapz_rows(avg_len, _ALPHA_STEPS)
    auc_abs = _trapz_rows(abs_err, _ALPHA_STEPS)
    auc_neg = _trapz_rows(neg_err, _ALPHA_STEPS)
    return [{
        "coverage_values": coverage[i],
        "avg_length_values": avg_len[i],
        "coverage_error_values": err[i],
        "abs_coverage_error_values": abs_err[i],
        "neg_coverage_error_values": neg_err[i],
        "auc_abs_error_values": auc_abs[i],
        "auc_length_values": auc_len[i],
        "auc_neg_error_values": auc_neg[i],
    } for i in range(hist.shape[0])]


def _auce_from_hist(hist: np.ndarray, sigma_sum: float, n: float, z: np.ndarray) -> Dict[str, object]:
    return _auce_from_hist_batch(np.asarray(hist)[None], np.array([sigma_sum], dtype=np.float64),
                                 np.array([n], dtype=np.float64), z)[0]


ArrayLike = Union[np.ndarray, Tensor]


def _to_cuda_f32(a: ArrayLike, device) -> Tensor:
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return a.to(device=device, dtype=torch.float32, non_blocking=True)


def auce(mean_values: ArrayLike, sigma_values: ArrayLike, target_values: ArrayLike, device=None) -> Dict[str, object]:
    """Drop-in for ``nerfuncertainty.metrics.auce`` (auce.py:10-57).  Accepts the reference's host numpy
    arrays (copied to ``device``, default the current CUDA device) or CUDA tensors; all three must share
    one shape (sigma per element)."""
    if device is None:
        device = mean_values.device if isinstance(mean_values, torch.Tensor) and mean_values.is_cuda \
            else torch.device("cuda", torch.cuda.current_device())
    m = _to_cuda_f32(mean_values, device).reshape(-1, 1)
    s = _to_cuda_f32(sigma_values, device).reshape(-1)
    t = _to_cuda_f32(target_values, device).reshape(-1, 1)
    if not (m.shape[0] == s.shape[0] == t.shape[0]):
        raise ValueError("mean_values, sigma_values and target_values must have the same shape")
    n = m.shape[0]
    z = _z_table(device)
    out = ops.score_prologue(m, t, s, [n], z, nll_min_std=0.0, sigma_from_var=False, want_vectors=False)
    hist = out["hist"][0].cpu().numpy()
    sigma_sum = float(out["sums"][0, 4].item())
    return _auce_from_hist(hist, sigma_sum, float(n), z_values_host())


def _score_rgb_device(rgb_pred: Tensor, rgb_gt: Tensor, rgb_std: Tensor, min_rgb_std_for_nll: float):
    """The device half of ``score_rgb_batch``: prologue (+ AUCE histogram, NLL) and AUSE slice sums, every kernel
    writing straight into its region of ONE packed float64 buffer (``B x 400`` curve sums | ``B x 5`` scalar sums |
    ``B x (nz + 1)`` histogram rows, int64 bit patterns) that the host tail reads back with a single copy.
    Enqueue-only (CUDA-graph capturable once the segment / cut tables of the shape are cached)."""
    if rgb_pred.dim() == 3:
        rgb_pred, rgb_gt, rgb_std = rgb_pred[None], rgb_gt[None], rgb_std[None]
    b, h, w, c = rgb_pred.shape
    n = h * w
    lens = [n] * b
    z = _z_table(rgb_pred.device)
    select = _use_select(n)
    nzp = z.numel() + 1
    packed_dev = torch.empty(b * (4 * N_RATIOS + 5 + nzp), dtype=torch.float64, device=rgb_pred.device)
    sums_v, psum_v, hist_v = _packed_views(packed_dev, b, nzp)
    pro = ops.score_prologue(rgb_pred.reshape(-1, c), rgb_gt.reshape(-1, c), rgb_std.reshape(-1), lens, z,
                             nll_min_std=min_rgb_std_for_nll, sigma_from_var=True, want_vectors=True,
                             want_coarse=select, out_sums=psum_v, out_hist=hist_v.view(torch.int64))
    vec = pro["vectors"]                                   # [3, total]: var, abs err, sq err
    cuts_one = ause_cut_counts(n)
    cuts = _tiled_cuts(n, b)
    _ause_sums(vec, lens, cuts, pro.get("coarse"), out=sums_v)                        # [B, 4, 100]
    return packed_dev, b, n, c, cuts_one


def _packed_views(packed, b: int, nzp: int):
    """The three regions of the packed score buffer (torch tensor or numpy array): ``[B, 4, 100]`` curve sums,
    ``[B, 5]`` scalar sums, ``[B, nz + 1]`` histogram rows (float64 storage holding int64 bit patterns)."""
    s0, s1 = b * 4 * N_RATIOS, b * (4 * N_RATIOS + 5)
    return packed[:s0].reshape(b, 4, N_RATIOS), packed[s0:s1].reshape(b, 5), packed[s1:].reshape(b, nzp)


_TILED: Dict[Tuple[int, int], np.ndarray] = {}


def _tiled_cuts(n: int, b: int) -> np.ndarray:
    key = (n, b)
    t = _TILED.get(key)
    if t is None:
        if len(_TILED) > 64:
            _TILED.clear()
        t = _TILED[key] = np.tile(ause_cut_counts(n)[None, :], (b, 1))
    return t


def score_rgb_batch_async(rgb_pred: Tensor, rgb_gt: Tensor, rgb_std: Tensor, min_rgb_std_for_nll: float = 3e-2
                          ) -> "PendingScores":
    """``get_unc_metrics_rgb`` (eval_uncertainty.py:306-402) for a batch of images in one set of
    segmented launches.  ``rgb_pred, rgb_gt [B, H, W, 3]``, ``rgb_std [B, H, W, 1]`` (CUDA float32; the
    ground truth already composited with the background for splat models).  Returns one dict per image
    with the reference's scalar / curve entries (``nll_rgb``, ``ause_*``, ``err_*``, ``err_var_*``,
    ``avg_var``, ``mse_mean`` and the 8 AUCE entries).

    Device work: 1 prologue (+ finalize), the AUSE slice sums (multi-cut select; or ONE segmented sort over 3B
    segments + 1 cut-point prefix-sum pass with ``UB_AUSE_SORT=1``), 1 packed device->host copy."""
    packed_dev, b, n, c, cuts_one = _score_rgb_device(rgb_pred, rgb_gt, rgb_std, min_rgb_std_for_nll)
    # one device->host transfer for everything the host tail needs (asynchronous into pinned memory)
    packed_host = torch.empty(packed_dev.shape, dtype=packed_dev.dtype, pin_memory=True)
    packed_host.copy_(packed_dev, non_blocking=True)
    done = torch.cuda.Event()
    done.record()
    return PendingScores(packed_host, packed_dev, done, b, n, c, cuts_one)


class PendingScores:
    """Device work of ``score_rgb_batch`` in flight: ``finish()`` waits for the packed result and runs the
    numpy tail.  Lets a driver enqueue the next view before the previous one's record is read back."""

    def __init__(self, packed_host, packed_dev, done, b, n, c, cuts_one):
        self.packed_host, self.packed_dev, self.done = packed_host, packed_dev, done
        self.b, self.n, self.c, self.cuts_one = b, n, c, cuts_one
        self._result: Optional[List[Dict[str, object]]] = None

    def finish(self) -> List[Dict[str, object]]:
        if self._result is None:
            self._result = self._finish()
        return self._result

    def _finish(self) -> List[Dict[str, object]]:
        self.done.synchronize()
        packed = self.packed_host.numpy()
        if os.environ.get("UB_NUMPY_TAIL", "0") != "1":
            return _native_tail(packed, self.b, self.n, self.c, self.cuts_one)
        return _numpy_tail(packed, self.b, self.n, self.c, self.cuts_one)


def _native_tail(packed: np.ndarray, b: int, n, c: int, cuts_one: np.ndarray, nll_key: str = "nll_rgb"
                 ) -> List[Dict[str, object]]:
    """The host tail of ``score_rgb_batch`` in one native call (``ub_score_tail_host``, csrc/score_tail.cu): the
    reference's numpy / torch expressions downstream of the device results, operation by operation in the reference's
    dtypes -- bit-identical to ``_numpy_tail`` (tests/test_host_logic.py), ~2 us per image instead of 0.2 ms + 17 us."""
    from . import _lib

    lib = _lib.load()
    zh = z_values_host()
    _alpha_tables()
    nz = len(zh)
    packed = np.ascontiguousarray(packed, dtype=np.float64)
    cuts = np.ascontiguousarray(cuts_one, dtype=np.int64)
    if cuts.ndim == 1:                                   # one image size for the whole batch (rgb)
        cuts = np.ascontiguousarray(np.broadcast_to(cuts, (b, cuts.shape[0])))
    nc = cuts.shape[1]
    npix = np.ascontiguousarray(np.broadcast_to(np.asarray(n, dtype=np.int64), (b,)))
    by_unc, o64 = np.empty((b, 3, nc)), np.empty((b, 3, nc))
    o32, is64 = np.empty((b, 3, nc), dtype=np.float32), np.empty((b, 3), dtype=np.int32)
    ause_v, scal = np.empty((b, 3)), np.empty((b, 3), dtype=np.float32)
    curves, auc = np.empty((b, 5, nz)), np.empty((b, 3))
    _lib.check(lib.ub_score_tail_host(packed.ctypes.data, b, npix.ctypes.data, c, cuts.ctypes.data, nc, _RATIO_STEPS.ctypes.data,
                                      zh.ctypes.data, nz, _ONE_MINUS_ALPHA.ctypes.data, _ALPHA_STEPS.ctypes.data,
                                      by_unc.ctypes.data, o64.ctypes.data, o32.ctypes.data, is64.ctypes.data,
                                      ause_v.ctypes.data, scal.ctypes.data, curves.ctypes.data, auc.ctypes.data))
    # rows as lists of array views / scalars (one C call each) instead of b x 20 numpy index expressions
    o64r, o32r, bur = list(o64.reshape(b * 3, nc)), list(o32.reshape(b * 3, nc)), list(by_unc.reshape(b * 3, nc))
    w64, au, sc = is64.ravel().tolist(), list(ause_v.ravel()), scal.ravel().tolist()
    cv, ac = list(curves.reshape(b * 5, nz)), list(auc.ravel())
    results = []
    for i in range(b):
        j, q = 3 * i, 5 * i
        results.append({
            "err_mae": o64r[j] if w64[j] else o32r[j], "err_var_mae": bur[j], "ause_mae": au[j],
            "err_mse": o64r[j + 1] if w64[j + 1] else o32r[j + 1], "err_var_mse": bur[j + 1], "ause_mse": au[j + 1],
            "err_rmse": o64r[j + 2] if w64[j + 2] else o32r[j + 2], "err_var_rmse": bur[j + 2], "ause_rmse": au[j + 2],
            nll_key: sc[j], "avg_var": sc[j + 1], "mse_mean": sc[j + 2],
            "coverage_values": cv[q], "avg_length_values": cv[q + 1], "coverage_error_values": cv[q + 2],
            "abs_coverage_error_values": cv[q + 3], "neg_coverage_error_values": cv[q + 4],
            "auc_abs_error_values": ac[j], "auc_length_values": ac[j + 1], "auc_neg_error_values": ac[j + 2],
        })
    return results


def _numpy_tail(packed: np.ndarray, b: int, n: int, c: int, cuts_one: np.ndarray) -> List[Dict[str, object]]:
    """The same tail as numpy expressions (the statement the native tail is checked against; ``UB_NUMPY_TAIL=1``
    routes ``finish()`` through it)."""
    zh = z_values_host()
    sums_v, psums, hist_v = _packed_views(packed, b, len(zh) + 1)
    bu_ae, bu_se, or_ae, or_se = sums_v[:, 0], sums_v[:, 1], sums_v[:, 2], sums_v[:, 3]
    hist = np.ascontiguousarray(hist_v).view(np.int64)
    # all twelve curves of the batch in three numpy expressions: rows = (image, {mae, mse, rmse})
    ora = _prefix_means(np.stack([or_ae, or_se, or_se], axis=1), cuts_one, "mae")      # [B, 3, 100] float32
    byu = _prefix_means(np.stack([bu_ae, bu_se, bu_se], axis=1), cuts_one, "mae")
    with np.errstate(invalid="ignore"):
        ora[:, 2] = np.sqrt(ora[:, 2])                                                   # rmse: torch.sqrt of the mean
        byu[:, 2] = np.sqrt(byu[:, 2])
    o_all, v_all, a_all = _ause_tail_batch(ora.reshape(b * 3, 100), byu.reshape(b * 3, 100))
    tails = {et: (o_all[k::3], v_all[k::3], a_all[k::3]) for k, et in enumerate(("mae", "mse", "rmse"))}
    nll = (psums[:, 3] / (n * c)).astype(np.float32)
    avg_var = (psums[:, 2] / n).astype(np.float32)
    mse_mean = (psums[:, 0] / n).astype(np.float32)
    auce_rows = _auce_from_hist_batch(hist, psums[:, 4] * c, np.full(b, float(n * c)), zh)
    results = []
    for i in range(b):
        d: Dict[str, object] = {}
        for et in ("mae", "mse", "rmse"):
            o, v, a = tails[et]
            d[f"err_{et}"], d[f"err_var_{et}"], d[f"ause_{et}"] = o[i], v[i], a[i]
        d["nll_rgb"] = float(nll[i])
        d["avg_var"] = float(avg_var[i])
        d["mse_mean"] = float(mse_mean[i])
        d.update(auce_rows[i])
        results.append(d)
    return results


def score_rgb_batch(rgb_pred: Tensor, rgb_gt: Tensor, rgb_std: Tensor, min_rgb_std_for_nll: float = 3e-2
                    ) -> List[Dict[str, object]]:
    """Synchronous form: ``score_rgb_batch_async(...).finish()``."""
    return score_rgb_batch_async(rgb_pred, rgb_gt, rgb_std, min_rgb_std_for_nll).finish()


def per_image_rgb_scalars(d: Dict[str, object]) -> Dict[str, float]:
    """The per-image scalar entries of ``metrics_dict`` (eval_uncertainty.py:765-776)."""
    mse = float(d["mse_mean"])
    return {
        "rgb_ause_mse": float(d["ause_mse"]), "rgb_ause_mae": float(d["ause_mae"]),
        "rgb_ause_rmse": float(d["ause_rmse"]), "rgb_mse": mse, "rgb_rmse": float(np.sqrt(mse)),
        "rgb_nll": float(d["nll_rgb"]), "rgb_avg_var": float(d["avg_var"]),
        "rgb_auc_abs_error": d["auc_abs_error_values"], "rgb_auc_length": d["auc_length_values"],
        "rgb_auc_neg_error": d["auc_neg_error_values"],
    }


def per_image_depth_scalars(d: Dict[str, object]) -> Dict[str, float]:
    """The per-image depth entries of ``metrics_dict`` (eval_uncertainty.py:714-725)."""
    mse = float(d["mse_mean"])
    return {
        "depth_ause_mse": float(d["ause_mse"]), "depth_ause_mae": float(d["ause_mae"]),
        "depth_ause_rmse": float(d["ause_rmse"]), "depth_mse": mse, "depth_rmse": float(np.sqrt(mse)),
        "depth_nll": float(d["nll_depth"]), "depth_avg_var": float(d["avg_var"]),
        "depth_auc_abs_error": d["auc_abs_error_values"], "depth_auc_length": d["auc_length_values"],
        "depth_auc_neg_error": d["auc_neg_error_values"],
    }


def resize_like_reference(img: Tensor, size) -> Tensor:
    """``[B, h, w] -> [B, H, W]`` the way ``F.resize(x.view(1, 1, h, w), size=size, antialias=None)`` does it per view
    (eval_uncertainty.py:444-451): torchvision's resize of a float tensor is ``interpolate(mode="bilinear",
    align_corners=False, antialias=False)``.  A data-format adapter in front of the hot path (plain torch)."""
    return torch.nn.functional.interpolate(img[:, None].float(), size=tuple(size), mode="bilinear", align_corners=False,
                                           antialias=False)[:, 0]


def score_depth_batch(depth: Tensor, depth_std: Tensor, depth_gt: Tensor, scales: Sequence[float],
                      min_depth_std_for_nll: float = 1.0) -> List[Dict[str, object]]:
    """``get_unc_metrics_depth`` (eval_uncertainty.py:415-644) for a batch of views, downstream of file
    loading.  ``depth, depth_std [B, h, w]``, ``depth_gt [B, H, W]`` (or with a trailing 1); renders whose shape
    differs from the ground truth's (splatfacto renders ``[H-1, W-1]`` depth) are resized like the reference does
    (``:442-452``: bilinear, ``align_corners=False``, no antialiasing).  ``scales[b]`` is the per-dataset scale ``a``.  Per view: scale, clamp to ``[1e-3, max gt]``, keep the
    pixels with ``gt > 0`` (a ragged segment per view), then the same kernels as the rgb path with one
    channel and sigma = std: prologue (se / ae / var, NLL with eps = ``min_depth_std_for_nll``, interval
    histogram), one segmented sort over 3B ragged segments, cut-point prefix sums.  The per-view preparation is
    one stable stream compaction (``ub_depth_prepare``); the per-view lengths are the only value read back."""
    if depth.dim() == 4:
        depth, depth_std = depth[..., 0], depth_std[..., 0]
    if depth_gt.dim() == 4:
        depth_gt = depth_gt[..., 0]
    if depth.shape[-2:] != depth_gt.shape[-2:]:
        depth = resize_like_reference(depth, depth_gt.shape[-2:])
    if depth_std.shape[-2:] != depth_gt.shape[-2:]:
        depth_std = resize_like_reference(depth_std, depth_gt.shape[-2:])
    b = depth.shape[0]
    dev = depth.device
    pred, std, gt, lens = ops.depth_prepare(depth.reshape(b, -1), depth_std.reshape(b, -1), depth_gt.reshape(b, -1),
                                            scales)
    pred, gt = pred.reshape(-1, 1), gt.reshape(-1, 1)
    total = pred.shape[0]
    z = _z_table(dev)
    select = _use_select(max(lens) if len(lens) else 0)
    nzp = z.numel() + 1
    packed_dev = torch.empty(b * (4 * N_RATIOS + 5 + nzp), dtype=torch.float64, device=dev)
    sums_v, psum_v, hist_v = _packed_views(packed_dev, b, nzp)
    pro = ops.score_prologue(pred, gt, std, lens, z, nll_min_std=min_depth_std_for_nll, sigma_from_var=False,
                             want_vectors=True, want_coarse=select, out_sums=psum_v, out_hist=hist_v.view(torch.int64))
    vec = pro["vectors"]
    cuts = np.stack([ause_cut_counts(n) for n in lens])
    _ause_sums(vec, lens, cuts, pro.get("coarse"), out=sums_v)
    packed = packed_dev.cpu().numpy()                      # one device->host copy of the three regions
    if os.environ.get("UB_NUMPY_TAIL", "0") != "1":
        return _native_tail(packed, b, np.asarray(lens, dtype=np.int64), 1, cuts, nll_key="nll_depth")
    return _numpy_depth_tail(packed, b, lens, cuts)


def _numpy_depth_tail(packed: np.ndarray, b: int, lens: Sequence[int], cuts: np.ndarray) -> List[Dict[str, object]]:
    """The depth tail as numpy expressions, one view at a time (ragged lengths); the statement the native tail is
    checked against."""
    zh = z_values_host()
    sums_a, psums_a, hist_a = _packed_views(packed, b, len(zh) + 1)
    hist_i = np.ascontiguousarray(hist_a).view(np.int64)
    results = []
    for i in range(b):
        n, ci = lens[i], cuts[i]
        bu_ae, bu_se, or_ae, or_se = sums_a[i, 0], sums_a[i, 1], sums_a[i, 2], sums_a[i, 3]
        psums, hist = psums_a[i], hist_i[i]
        d: Dict[str, object] = {}
        _, d["err_mae"], d["err_var_mae"], d["ause_mae"] = _ause_tail(
            _prefix_means(or_ae, ci, "mae"), _prefix_means(bu_ae, ci, "mae"))
        _, d["err_mse"], d["err_var_mse"], d["ause_mse"] = _ause_tail(
            _prefix_means(or_se, ci, "mse"), _prefix_means(bu_se, ci, "mse"))
        _, d["err_rmse"], d["err_var_rmse"], d["ause_rmse"] = _ause_tail(
            _prefix_means(or_se, ci, "rmse"), _prefix_means(bu_se, ci, "rmse"))
        with np.errstate(divide="ignore", invalid="ignore"):
            d["nll_depth"] = float(np.float32(psums[3] / n)) if n else float("nan")
            d["avg_var"] = float(np.float32(psums[2] / n)) if n else float("nan")
            d["mse_mean"] = float(np.float32(psums[0] / n)) if n else float("nan")
        d.update(_auce_from_hist(hist, float(psums[4]), float(n), zh))
        results.append(d)
    return results
