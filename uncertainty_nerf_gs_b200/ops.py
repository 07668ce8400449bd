"""Torch-tensor front end of the C ABI: shape checks, output / workspace allocation, stream passing.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in ``libub200.so``.
Every function requires CUDA float32 tensors and raises otherwise -- there is no CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import UBError  # noqa: F401  (re-exported)

Tensor = torch.Tensor

# kernels launched by this process through the C ABI (bench.py reports it as `gpu_launches`)
LAUNCH_COUNT = 0


def _count(n: int) -> None:
    global LAUNCH_COUNT
    LAUNCH_COUNT += n


def _dev_f32(t: Tensor, name: str) -> Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: the ub200 path has no CPU implementation")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_NO_GUARD = contextlib.nullcontext()


def _stream() -> int:
    """Raw ``cudaStream_t`` of torch's current stream on the current device (the fast private accessor when this
    torch has it: ``torch.cuda.current_stream()`` costs ~20 us of Python per call)."""
    if _RAW_STREAM is not None:
        return _RAW_STREAM(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _guard(device):
    """``torch.cuda.device(device)`` only when the current device differs (the context manager itself costs ~10 us)."""
    idx = device.index if isinstance(device, torch.device) else int(device)
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _squeeze_last(t: Tensor, name: str, ndim: int) -> Tensor:
    if t.dim() == ndim + 1 and t.shape[-1] == 1:
        t = t[..., 0]
    if t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims (or a trailing singleton), got shape {tuple(t.shape)}")
    return t


_BG_MODES = {"last_sample": _lib.UB_BG_LAST_SAMPLE, "random": _lib.UB_BG_NONE, "none": _lib.UB_BG_NONE}
_BETA_MODES = {"raw": _lib.UB_BETA_RAW, "nan_guard": _lib.UB_BETA_NAN_GUARD}


def _ray_stream(t: Tensor, name: str, R: int, S: int, C_: int = 1) -> Tensor:
    """A ``[R, S]`` / ``[R, S, 1]`` (``C_ == 1``) or ``[R, S, C_]`` float32 CUDA stream, contiguous; only the
    pointer is used afterwards, so no reshaped view is created (each costs microseconds on this latency-bound path)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: the ub200 path has no CPU implementation")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    shp = t.shape
    ok = (len(shp) == 3 and shp[0] == R and shp[1] == S and shp[2] == C_) or \
         (C_ == 1 and len(shp) == 2 and shp[0] == R and shp[1] == S)
    if not ok:
        want = f"{(R, S, C_)}" if C_ != 1 else f"{(R, S)} or {(R, S, 1)}"
        raise ValueError(f"{name}: shape {tuple(shp)} != {want}")
    return t if t.is_contiguous() else t.contiguous()


def composite_rays(density: Tensor, deltas: Optional[Tensor], starts: Tensor, ends: Tensor, rgb: Tensor,
                   beta: Optional[Tensor] = None, *, background: Union[str, Sequence[float], Tensor] = "last_sample",
                   beta_mode: str = "nan_guard", rays_per_chunk: Optional[int] = None,
                   eval_mode: bool = True, return_weights: bool = False,
                   image_hw: Optional[Tuple[int, int]] = None) -> Dict[str, Tensor]:
    """One fused pass over ``[R, S]`` ray samples -> rgb, accumulation, median depth, expected depth,
    ``rgb_var = sum w^2 beta``, depth variance (+ std of both).  Outputs are ``[R, C]`` like the
    reference's per-chunk ``get_outputs`` (activenerfacto_model.py:94-127), or ``[H, W, C]`` when
    ``image_hw = (H, W)`` with ``H * W == R`` (what ``get_outputs_for_camera`` makes of them).
    ``deltas=None``: the bin widths are ``ends - starts`` (float32) -- bit for bit what nerfstudio's
    ``RayBundle.get_ray_samples`` stores in ``RaySamples.deltas`` -- and the kernel reads one stream less (1384 instead
    of 1576 B per 48-sample ray)."""
    lib = _lib.load()
    if not isinstance(density, torch.Tensor) or density.dim() not in (2, 3):
        raise ValueError("density: expected [R, S] or [R, S, 1]")
    R, S = int(density.shape[0]), int(density.shape[1])
    density = _ray_stream(density, "density", R, S)
    if deltas is not None:
        deltas = _ray_stream(deltas, "deltas", R, S)
    starts = _ray_stream(starts, "starts", R, S)
    ends = _ray_stream(ends, "ends", R, S)
    rgb = _ray_stream(rgb, "rgb", R, S, 3)
    if beta is not None:
        beta = _ray_stream(beta, "beta", R, S)
    dev = density.device
    args = _lib.CompositeRaysArgs()
    args.density, args.deltas = density.data_ptr(), _ptr(deltas)
    args.starts, args.ends = starts.data_ptr(), ends.data_ptr()
    args.rgb, args.beta = rgb.data_ptr(), _ptr(beta)
    args.num_rays, args.num_samples = R, S
    if isinstance(background, str):
        if background not in _BG_MODES:
            raise ValueError(f"unsupported background {background!r}")
        args.background_mode = _BG_MODES[background]
    else:
        bg = [float(v) for v in (background.tolist() if isinstance(background, torch.Tensor) else background)]
        if len(bg) != 3:
            raise ValueError("fixed background must have 3 components")
        args.background_mode = _lib.UB_BG_FIXED
        args.background_rgb = (C.c_float * 3)(*bg)
    args.beta_mode = _BETA_MODES[beta_mode]
    args.rays_per_chunk = int(rays_per_chunk) if rays_per_chunk else 0
    args.eval_mode = 1 if eval_mode else 0
    # one allocation for all per-ray outputs (rows of a [10, R] buffer + the chunk workspace behind it):
    # the wrapper, not the kernel, bounds the latency of small (training-sized) batches
    ws_bytes = lib.ub_composite_rays_workspace_bytes(R, args.rays_per_chunk)
    ws_floats = (ws_bytes + 3) // 4
    buf = torch.empty(10 * R + ws_floats + 4, device=dev)
    base = buf.data_ptr()
    if image_hw is not None:
        if image_hw[0] * image_hw[1] != R:
            raise ValueError(f"image_hw {tuple(image_hw)} does not hold {R} rays")
        lead = (int(image_hw[0]), int(image_hw[1]))
    else:
        lead = (R,)
    rows = buf[3 * R:10 * R].view(7, *lead, 1).unbind(0)     # accumulation, depth, expected, var, std, dvar, dstd
    out = {"rgb": buf[:3 * R].view(*lead, 3), "accumulation": rows[0], "depth": rows[1], "expected_depth": rows[2],
           "depth_var": rows[5], "depth_std": rows[6]}
    f = 4 * R
    args.out_rgb, args.out_accumulation = base, base + 3 * f
    args.out_depth, args.out_expected_depth = base + 4 * f, base + 5 * f
    args.out_depth_var, args.out_depth_std = base + 8 * f, base + 9 * f
    if beta is not None:
        out["rgb_var"], out["rgb_std"] = rows[3], rows[4]
        args.out_rgb_var, args.out_rgb_std = base + 6 * f, base + 7 * f
    if return_weights:
        out["weights"] = torch.empty(R, S, 1, device=dev)
        args.out_weights = out["weights"].data_ptr()
    ws_ptr = base + 10 * f
    ws_ptr += (-ws_ptr) % 16
    ws_off = (ws_ptr - base) // 4
    out["_workspace"] = buf[ws_off:ws_off + ws_floats]   # chunk clip bounds (used by the backward)
    with _guard(dev):
        _lib.check(lib.ub_composite_rays(C.byref(args), ws_ptr, ws_bytes, _stream()))
    _count(2 if R > 0 else 0)
    return out


def composite_rays_many(batches: Sequence[Sequence[Tensor]], *, background: Union[str, Sequence[float]] = "last_sample",
                        beta_mode: str = "nan_guard", rays_per_chunk: Optional[int] = None, eval_mode: bool = True,
                        image_hw: Optional[Tuple[int, int]] = None) -> List[Dict[str, Tensor]]:
    """``composite_rays`` for several independent ray batches of one shape -- the M members of a view -- through
    ``ub_composite_rays_batch``: one workspace memset, the compositing kernels back to back, one finalize launch,
    one host call.  ``batches[i] = (density, deltas, starts, ends, rgb, beta)``; returns one output dict per batch."""
    lib = _lib.load()
    n = len(batches)
    if n < 1:
        return []
    if n > _lib.UB_MAX_COMPOSITE_BATCH:
        out: List[Dict[str, Tensor]] = []
        for lo in range(0, n, _lib.UB_MAX_COMPOSITE_BATCH):
            out += composite_rays_many(batches[lo:lo + _lib.UB_MAX_COMPOSITE_BATCH], background=background,
                                       beta_mode=beta_mode, rays_per_chunk=rays_per_chunk, eval_mode=eval_mode,
                                       image_hw=image_hw)
        return out
    first = batches[0][0]
    if not isinstance(first, torch.Tensor) or first.dim() not in (2, 3):
        raise ValueError("density: expected [R, S] or [R, S, 1]")
    R, S = int(first.shape[0]), int(first.shape[1])
    dev = first.device
    if isinstance(background, str):
        if background not in _BG_MODES:
            raise ValueError(f"unsupported background {background!r}")
        bg_mode, bg = _BG_MODES[background], None
    else:
        bg = [float(v) for v in background]
        if len(bg) != 3:
            raise ValueError("fixed background must have 3 components")
        bg_mode = _lib.UB_BG_FIXED
    if image_hw is not None:
        if image_hw[0] * image_hw[1] != R:
            raise ValueError(f"image_hw {tuple(image_hw)} does not hold {R} rays")
        lead = (int(image_hw[0]), int(image_hw[1]))
    else:
        lead = (R,)
    arr = (_lib.CompositeRaysArgs * n)()
    keep = []
    for a, batch in zip(arr, batches):
        density, deltas, starts, ends, rgb, beta = batch
        density = _ray_stream(density, "density", R, S)
        if deltas is not None:
            deltas = _ray_stream(deltas, "deltas", R, S)
        starts = _ray_stream(starts, "starts", R, S)
        ends = _ray_stream(ends, "ends", R, S)
        rgb = _ray_stream(rgb, "rgb", R, S, 3)
        beta = _ray_stream(beta, "beta", R, S)
        keep.append((density, deltas, starts, ends, rgb, beta))
        a.density, a.deltas, a.starts, a.ends = density.data_ptr(), _ptr(deltas), starts.data_ptr(), ends.data_ptr()
        a.rgb, a.beta = rgb.data_ptr(), beta.data_ptr()
        a.num_rays, a.num_samples = R, S
        a.background_mode = bg_mode
        if bg is not None:
            a.background_rgb = (C.c_float * 3)(*bg)
        a.beta_mode = _BETA_MODES[beta_mode]
        a.rays_per_chunk = int(rays_per_chunk) if rays_per_chunk else 0
        a.eval_mode = 1 if eval_mode else 0
    ws_bytes = lib.ub_composite_rays_batch_workspace_bytes(arr, n)
    ws_floats = (ws_bytes + 3) // 4
    per = 10 * R
    buf = torch.empty(n * per + ws_floats + 4, device=dev)
    base = buf.data_ptr()
    f = 4 * R
    outs = []
    for i, a in enumerate(arr):
        b0 = base + 4 * per * i
        a.out_rgb, a.out_accumulation = b0, b0 + 3 * f
        a.out_depth, a.out_expected_depth = b0 + 4 * f, b0 + 5 * f
        a.out_rgb_var, a.out_rgb_std = b0 + 6 * f, b0 + 7 * f
        a.out_depth_var, a.out_depth_std = b0 + 8 * f, b0 + 9 * f
        rows = buf[per * i + 3 * R:per * (i + 1)].view(7, *lead, 1).unbind(0)
        outs.append({"rgb": buf[per * i:per * i + 3 * R].view(*lead, 3), "accumulation": rows[0], "depth": rows[1],
                     "expected_depth": rows[2], "rgb_var": rows[3], "rgb_std": rows[4], "depth_var": rows[5],
                     "depth_std": rows[6]})
    ws_ptr = base + 4 * per * n
    ws_ptr += (-ws_ptr) % 16
    with _guard(dev):
        _lib.check(lib.ub_composite_rays_batch(arr, n, ws_ptr, ws_bytes, _stream()))
    _count(n + 1 if R > 0 else 0)
    return outs


def render_weights(weights: Tensor, starts: Tensor, ends: Tensor, *, rays_per_chunk: Optional[int] = None,
                   want: Sequence[str] = ("depth",)) -> Dict[str, Tensor]:
    """Median depth / expected depth / accumulation / depth variance from given weights
    (``prop_depth_i``, activenerfacto_model.py:150-151; averaged sampled weights, laplace_model.py:509-521)."""
    lib = _lib.load()
    weights = _dev_f32(_squeeze_last(weights, "weights", 2), "weights")
    R, S = weights.shape
    starts = _dev_f32(_squeeze_last(starts, "starts", 2), "starts")
    ends = _dev_f32(_squeeze_last(ends, "ends", 2), "ends")
    if starts.shape != (R, S) or ends.shape != (R, S):
        raise ValueError("starts / ends must match weights")
    dev = weights.device
    args = _lib.RenderWeightsArgs()
    args.weights, args.starts, args.ends = weights.data_ptr(), starts.data_ptr(), ends.data_ptr()
    args.num_rays, args.num_samples = R, S
    args.rays_per_chunk = int(rays_per_chunk) if rays_per_chunk else 0
    fields = {"accumulation": "out_accumulation", "depth": "out_depth", "expected_depth": "out_expected_depth",
              "depth_var": "out_depth_var", "depth_std": "out_depth_std"}
    out: Dict[str, Tensor] = {}
    for k in want:
        if k not in fields:
            raise ValueError(f"unknown output {k!r}")
        out[k] = torch.empty(R, 1, device=dev)
        setattr(args, fields[k], out[k].data_ptr())
    ws = _workspace(lib.ub_render_weights_workspace_bytes(R, args.rays_per_chunk), dev)
    with _guard(dev):
        _lib.check(lib.ub_render_weights(C.byref(args), ws.data_ptr(), ws.numel(), _stream()))
    _count((2 if "expected_depth" in out else 1) if R > 0 else 0)
    return out


def average_sampled_weights(density: Tensor, density_var: Tensor, deltas: Tensor, num_draws: int = 100,
                            noise: Optional[Tensor] = None, seed: int = 0) -> Tensor:
    """Mean weights of ``num_draws`` draws ``relu(N(density, sqrt(density_var)))`` (laplace_model.py:486-507)
    -> ``[R, S, 1]``.  ``noise [num_draws, R, S]`` (standard normal) makes the result comparable with the
    oracle; without it the draws come from an in-kernel Philox generator."""
    lib = _lib.load()
    density = _dev_f32(_squeeze_last(density, "density", 2), "density")
    R, S = density.shape
    density_var = _dev_f32(_squeeze_last(density_var, "density_var", 2), "density_var")
    deltas = _dev_f32(_squeeze_last(deltas, "deltas", 2), "deltas")
    if density_var.shape != (R, S) or deltas.shape != (R, S):
        raise ValueError("density_var / deltas must match density")
    if noise is not None:
        noise = _dev_f32(noise.reshape(-1, R, S), "noise")
        if noise.shape[0] != num_draws:
            raise ValueError("noise must be [num_draws, R, S]")
    out = torch.empty(R, S, 1, device=density.device)
    with _guard(density.device):
        _lib.check(lib.ub_average_sampled_weights(density.data_ptr(), density_var.data_ptr(), deltas.data_ptr(),
                                                  _ptr(noise), R, S, int(num_draws), int(seed), out.data_ptr(),
                                                  _stream()))
    _count(1 if R > 0 else 0)
    return out


_SPREAD = {None: _lib.UB_SPREAD_NONE, "std": _lib.UB_SPREAD_STD, "var": _lib.UB_SPREAD_VAR}


def reduce_many(jobs: Sequence[Tuple[Sequence[Tensor], Optional[str]]]) -> List[Tuple[Tensor, Optional[Tensor]]]:
    """Batched member reduce: ``jobs[j] = (members_j, spread_j)``; every job reduces K same-shaped
    ``[..., C]`` tensors (read in place, no stack) to their mean and, optionally, the unbiased std / var over
    members averaged over the channel axis (``[..., 1]``).  All jobs go through one kernel launch
    (chunks of ``UB_MAX_REDUCE_JOBS``)."""
    lib = _lib.load()
    if not jobs:
        return []
    k = len(jobs[0][0])
    prepared = []
    for j, (members, spread) in enumerate(jobs):
        if len(members) != k or k < 1:
            raise ValueError("every job must have the same, positive number of members")
        if spread not in _SPREAD:
            raise ValueError(f"spread must be None, 'std' or 'var', got {spread!r}")
        ms = [_dev_f32(m, f"jobs[{j}].members[{i}]") for i, m in enumerate(members)]
        shape = ms[0].shape
        if any(m.shape != shape for m in ms):
            raise ValueError("all members of a job must share one shape")
        if spread is None:
            c, n = 1, ms[0].numel()          # mean only: a flat stream
        else:
            c = int(shape[-1]) if len(shape) else 1
            n = ms[0].numel() // max(c, 1)
        dev = ms[0].device
        mean = torch.empty(shape, device=dev)
        spr = torch.empty((*shape[:-1], 1), device=dev) if spread else None
        prepared.append((ms, c, n, spread, mean, spr))
    dev = prepared[0][0][0].device
    results = []
    with _guard(dev):
        for lo in range(0, len(prepared), _lib.UB_MAX_REDUCE_JOBS):
            chunk = prepared[lo:lo + _lib.UB_MAX_REDUCE_JOBS]
            arr = (_lib.ReduceJob * len(chunk))()
            keep = []
            for slot, (ms, c, n, spread, mean, spr) in zip(arr, chunk):
                ptrs = (C.c_void_p * k)(*[m.data_ptr() for m in ms])
                keep.append(ptrs)
                slot.members_host = ptrs
                slot.num_pixels, slot.channels, slot.spread_mode = n, c, _SPREAD[spread]
                slot.out_mean, slot.out_spread = mean.data_ptr(), _ptr(spr)
            _lib.check(lib.ub_reduce_members_batched(arr, len(chunk), k, _stream()))
            _count(2)  # flat-mean and spread instantiations
    for (_, _, _, _, mean, spr) in prepared:
        results.append((mean, spr))
    return results


def reduce_members(members: Sequence[Tensor], spread: Optional[str] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """Mean over K same-shaped ``[..., C]`` member tensors and, optionally, the unbiased std / var over
    members averaged over the channel axis -> ``[..., 1]`` (single-job form of ``reduce_many``)."""
    if len(members) < 1:
        raise ValueError("need at least one member")
    return reduce_many([(members, spread)])[0]


def depth_prepare(depth: Tensor, depth_std: Tensor, depth_gt: Tensor, scales: Union[Sequence[float], Tensor]
                  ) -> Tuple[Tensor, Tensor, Tensor, List[int]]:
    """Scale, clamp and mask the depth-scoring inputs of a batch of views (eval_uncertainty.py:461-462, 511-513,
    552-560): ``depth, depth_std, depth_gt [B, N]`` -> the kept pixels as flat ``pred, std, gt`` (views back to
    back, row-major inside a view) and the per-view lengths.  Reading the lengths back is the one host
    synchronisation of the depth path."""
    lib = _lib.load()
    depth, depth_std, depth_gt = _dev_f32(depth, "depth"), _dev_f32(depth_std, "depth_std"), _dev_f32(depth_gt, "depth_gt")
    if depth.dim() != 2 or depth_std.shape != depth.shape or depth_gt.shape != depth.shape:
        raise ValueError("depth, depth_std, depth_gt must share one [B, N] shape")
    b, n = int(depth.shape[0]), int(depth.shape[1])
    dev = depth.device
    sc = scales if isinstance(scales, torch.Tensor) else torch.tensor([float(v) for v in scales], dtype=torch.float32)
    sc = sc.to(device=dev, dtype=torch.float32).contiguous()
    if sc.numel() != b:
        raise ValueError("one scale per view")
    pred, std, gt = (torch.empty(b * n, device=dev) for _ in range(3))
    offsets = torch.empty(b + 1, dtype=torch.int64, device=dev)
    ws = _workspace(lib.ub_depth_prepare_workspace_bytes(b, n), dev)
    with _guard(dev):
        _lib.check(lib.ub_depth_prepare(depth.data_ptr(), depth_std.data_ptr(), depth_gt.data_ptr(), sc.data_ptr(), b, n,
                                        pred.data_ptr(), std.data_ptr(), gt.data_ptr(), offsets.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _stream()))
    _count(4)
    off = offsets.cpu().tolist()
    total = off[-1]
    lens = [off[i + 1] - off[i] for i in range(b)]
    return pred[:total], std[:total], gt[:total], lens


class Segments:
    """Segment table of a batch (device int64 offsets + the host-side scalars the C ABI wants).  Built once
    per distinct tuple of lengths and cached: the hot calls never copy host memory to the device (a copy
    from pageable memory would synchronise the stream)."""

    _cache: Dict[tuple, "Segments"] = {}

    def __init__(self, seg_lengths: Sequence[int], device):
        lens = np.asarray(list(seg_lengths), dtype=np.int64)
        if (lens < 0).any():
            raise ValueError("segment lengths must be non-negative")
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        self.lengths = lens
        self.num = len(lens)
        self.total = int(off[-1])
        self.max_len = int(lens.max()) if len(lens) else 0
        self.offsets_host = off
        self.offsets = torch.from_numpy(off).to(device)
        self._cuts: Dict[bytes, Tensor] = {}

    @classmethod
    def get(cls, seg_lengths, device) -> "Segments":
        if isinstance(seg_lengths, Segments):
            return seg_lengths
        key = (str(device), tuple(int(v) for v in seg_lengths))
        seg = cls._cache.get(key)
        if seg is None:
            if len(cls._cache) > 64:
                cls._cache.clear()
            seg = cls._cache[key] = Segments(key[1], device)
        return seg

    def cuts_device(self, cuts: np.ndarray) -> Tensor:
        """Validated, cached device copy of a ``[num_segments, num_cuts]`` int64 cut table."""
        cuts = np.ascontiguousarray(cuts, dtype=np.int64)
        if cuts.ndim != 2 or cuts.shape[0] != self.num:
            raise ValueError("cuts must be [num_segments, num_cuts]")
        key = cuts.tobytes()
        t = self._cuts.get(key)
        if t is None:
            if (cuts < 0).any() or (cuts > self.lengths[:, None]).any():
                raise ValueError("every cut must lie in [0, segment length]")
            if len(self._cuts) > 16:
                self._cuts.clear()
            t = self._cuts[key] = torch.from_numpy(cuts).to(self.offsets.device)
        return t


def _own_or_new(t: Optional[Tensor], shape, dtype, device, name: str) -> Tensor:
    if t is None:
        return torch.empty(*shape, dtype=dtype, device=device)
    if tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != device or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous {dtype} tensor of shape {tuple(shape)} on {device}")
    return t


def score_prologue(pred: Tensor, target: Tensor, std: Tensor, seg_lengths: Sequence[int], z_values: Tensor,
                   nll_min_std: float, sigma_from_var: bool = True, want_vectors: bool = True,
                   want_coarse: bool = False, out_sums: Optional[Tensor] = None,
                   out_hist: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """se / ae / var vectors, float64 sums and the AUCE interval histogram for a batch of images.
    ``pred, target [N, C]``, ``std [N]``; images are consecutive segments of ``seg_lengths`` pixels.
    ``out_sums [nseg, 5]`` float64 / ``out_hist [nseg, nz + 1]`` int64: caller-owned contiguous outputs (e.g. slices
    of one packed result buffer) instead of fresh tensors.
    ``want_coarse``: also ``coarse [3, nseg, 4096]`` uint32 (as int32 storage), the top-12-bit key histograms of the
    three vectors that ``cut_select_sums(..., coarse=...)`` would otherwise compute with a pass of its own."""
    lib = _lib.load()
    pred = _dev_f32(pred, "pred")
    target = _dev_f32(target, "target")
    if pred.dim() != 2 or pred.shape != target.shape or pred.shape[1] not in (1, 3):
        raise ValueError("pred / target must be [N, 1] or [N, 3] and match")
    n, c = pred.shape
    std = _dev_f32(std.reshape(-1), "std")
    if std.numel() != n:
        raise ValueError("std must hold one value per pixel")
    seg = Segments.get(seg_lengths, pred.device)
    if seg.total != n:
        raise ValueError("segment lengths do not sum to the number of pixels")
    if z_values.dtype != torch.float64 or not z_values.is_cuda:
        raise TypeError("z_values must be a CUDA float64 tensor")
    dev = pred.device
    nseg, nz = seg.num, z_values.numel()
    out: Dict[str, Tensor] = {
        "sums": _own_or_new(out_sums, (nseg, _lib.UB_PROLOGUE_NSUMS), torch.float64, dev, "out_sums"),
        "hist": _own_or_new(out_hist, (nseg, nz + 1), torch.int64, dev, "out_hist"),
    }
    if want_vectors:
        # one [3, N] buffer (var, abs err, sq err) so that the three AUSE sorts run as one segmented sort
        out["vectors"] = torch.empty(3, n, device=dev)
        out["var"], out["absolute_error"], out["squared_error"] = out["vectors"][0], out["vectors"][1], out["vectors"][2]
    args = _lib.ScorePrologueArgs()
    args.pred, args.target, args.std = pred.data_ptr(), target.data_ptr(), std.data_ptr()
    args.channels, args.num_segments = c, nseg
    args.seg_offsets, args.max_segment_len = seg.offsets.data_ptr(), seg.max_len
    args.nll_min_std, args.sigma_from_var = float(nll_min_std), 1 if sigma_from_var else 0
    args.z_values, args.num_z = z_values.data_ptr(), nz
    args.out_sq_err, args.out_abs_err = _ptr(out.get("squared_error")), _ptr(out.get("absolute_error"))
    args.out_var = _ptr(out.get("var"))
    args.out_sums, args.out_hist = out["sums"].data_ptr(), out["hist"].data_ptr()
    if want_coarse:
        if not want_vectors:
            raise ValueError("want_coarse needs want_vectors")
        out["coarse"] = torch.empty(3, nseg, 4096, dtype=torch.int32, device=dev)
        args.out_coarse_hist = out["coarse"].data_ptr()
    ws = _workspace(lib.ub_score_prologue_workspace_bytes(nseg, seg.max_len, nz), dev)
    with _guard(dev):
        _lib.check(lib.ub_score_prologue(C.byref(args), ws.data_ptr(), ws.numel(), _stream()))
    _count(2)  # ratio table, prologue (its last block per segment folds the partial sums)
    return out


def segmented_sort(keys: Tensor, seg_lengths: Sequence[int], want_perm: bool = True, want_keys: bool = True
                   ) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """Stable ascending sort of each segment (``torch.sort(stable=True)`` order).  Returns
    ``(sorted_keys, perm)`` with ``perm`` int32 indices *within the segment*."""
    lib = _lib.load()
    keys = _dev_f32(keys.reshape(-1), "keys")
    seg = Segments.get(seg_lengths, keys.device)
    if seg.total != keys.numel():
        raise ValueError("segment lengths do not sum to the number of keys")
    dev = keys.device
    total = keys.numel()
    sorted_keys = torch.empty(total, device=dev) if want_keys else None
    perm = torch.empty(total, dtype=torch.int32, device=dev) if want_perm else None
    nseg, max_len = seg.num, seg.max_len
    ws = _workspace(lib.ub_segmented_sort_workspace_bytes(nseg, total, max_len, 1 if want_perm else 0), dev)
    with _guard(dev):
        _lib.check(lib.ub_segmented_sort(keys.data_ptr(), nseg, seg.offsets.data_ptr(), total, max_len,
                                         _ptr(sorted_keys), _ptr(perm), ws.data_ptr(), ws.numel(), _stream()))
    _count(12 if total > 0 else 0)
    return sorted_keys, perm


def cut_prefix_sums(values: Sequence[Tensor], perms: Union[None, Tensor, Sequence[Optional[Tensor]]],
                    seg_lengths: Sequence[int], cuts: np.ndarray) -> Tensor:
    """float64 ``sum_{i < cut} values_v[perm_v[i]]`` per segment, value array and cut -> ``[nseg, V, ncuts]``.
    ``perms``: one int32 permutation (within-segment indices, as ``segmented_sort`` returns) shared by all
    value arrays, or one per value array (``None`` = identity)."""
    lib = _lib.load()
    vals = [_dev_f32(v.reshape(-1), f"values[{i}]") for i, v in enumerate(values)]
    seg = Segments.get(seg_lengths, vals[0].device)
    total = seg.total
    if any(v.numel() != total for v in vals):
        raise ValueError("every value array must hold one entry per element")
    if perms is None or isinstance(perms, torch.Tensor):
        perms = [perms] * len(vals)
    if len(perms) != len(vals):
        raise ValueError("need one permutation (or None) per value array")
    for pm in perms:
        if pm is not None and (pm.dtype != torch.int32 or not pm.is_cuda or pm.numel() != total):
            raise TypeError("perm must be a CUDA int32 tensor with one entry per element")
    cuts_dev = seg.cuts_device(cuts)
    nseg, ncuts = seg.num, cuts_dev.shape[1]
    dev = vals[0].device
    out = torch.empty(nseg, len(vals), ncuts, dtype=torch.float64, device=dev)
    ptrs = (C.c_void_p * len(vals))(*[v.data_ptr() for v in vals])
    pptrs = (C.c_void_p * len(vals))(*[_ptr(pm) for pm in perms])
    ws = _workspace(lib.ub_cut_prefix_sums_workspace_bytes(nseg, seg.max_len, len(vals), ncuts), dev)
    with _guard(dev):
        _lib.check(lib.ub_cut_prefix_sums(ptrs, pptrs, len(vals), nseg, seg.offsets.data_ptr(), seg.max_len,
                                          cuts_dev.data_ptr(), ncuts, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                          _stream()))
    _count(2)
    return out


def cut_select_sums(families: Sequence[Tuple[Tensor, Tensor, Optional[Tensor]]], seg_lengths: Sequence[int],
                    cuts: np.ndarray, coarse: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """float64 sums of the payloads over the first ``cuts[s, c]`` elements of the stable ascending order of
    the keys, without sorting (``ub_cut_select_sums``): equal to ``segmented_sort`` + ``cut_prefix_sums`` up
    to float64 summation order.  ``families``: ``(keys, payload0, payload1 or None)`` triples of ``[total]``
    float32 tensors sharing the segmentation; ``payload0 is keys`` (same storage) sums the sorted keys.
    Returns ``[nseg, V, ncuts]`` with one row per payload array in family order.
    ``coarse [num_families, nseg, 4096]`` (int32 storage): the top-12-bit key histograms, when the producer of the
    keys already has them (``score_prologue(want_coarse=True)``).  ``out``: caller-owned contiguous
    ``[nseg, V, ncuts]`` float64 output."""
    lib = _lib.load()
    fams = []
    for i, (k, p0, p1) in enumerate(families):
        k = _dev_f32(k.reshape(-1), f"keys[{i}]")
        p0 = k if p0 is None or p0.data_ptr() == k.data_ptr() else _dev_f32(p0.reshape(-1), f"payload0[{i}]")
        p1 = None if p1 is None else _dev_f32(p1.reshape(-1), f"payload1[{i}]")
        fams.append((k, p0, p1))
    dev = fams[0][0].device
    seg = Segments.get(seg_lengths, dev)
    total = seg.total
    for k, p0, p1 in fams:
        if k.numel() != total or p0.numel() != total or (p1 is not None and p1.numel() != total):
            raise ValueError("every key / payload array must hold one entry per element")
    cuts_dev = seg.cuts_device(cuts)
    nseg, ncuts, nf = seg.num, cuts_dev.shape[1], len(fams)
    nrows = sum(1 if p1 is None else 2 for _, _, p1 in fams)
    out = _own_or_new(out, (nseg, nrows, ncuts), torch.float64, dev, "out")
    kp = (C.c_void_p * nf)(*[k.data_ptr() for k, _, _ in fams])
    p0p = (C.c_void_p * nf)(*[p0.data_ptr() for _, p0, _ in fams])
    p1p = (C.c_void_p * nf)(*[_ptr(p1) for _, _, p1 in fams])
    if coarse is not None:
        if coarse.dtype != torch.int32 or not coarse.is_cuda or tuple(coarse.shape) != (nf, nseg, 4096):
            raise ValueError(f"coarse must be a CUDA int32 tensor of shape {(nf, nseg, 4096)}")
        coarse = coarse.contiguous()
    ws = _workspace(lib.ub_cut_select_sums_workspace_bytes(nf, nseg, total, seg.max_len, ncuts), dev)
    with _guard(dev):
        _lib.check(lib.ub_cut_select_sums_ex(kp, p0p, p1p, nf, nseg, seg.offsets.data_ptr(), total, seg.max_len,
                                             cuts_dev.data_ptr(), ncuts, _ptr(coarse), out.data_ptr(), ws.data_ptr(),
                                             ws.numel(), _stream()))
    _count(7 if coarse is None else 6)  # [coarse], alloc, fine, locate, classify, resolve, finish
    return out


_ACTS = {"identity": _lib.UB_ACT_IDENTITY, "sigmoid": _lib.UB_ACT_SIGMOID, "exp": _lib.UB_ACT_EXP,
         "trunc_exp": _lib.UB_ACT_EXP}


def laplace_ll_moments(x: Tensor, sampled_params: Tensor, out_dim: int, activation: str,
                       want_mean2: bool = False, tensor_cores: bool = True) -> Dict[str, Tensor]:
    """E[y], E[y^2]-E[y]^2 of ``y = act(x W_s^T + b_s)`` over the rows of ``sampled_params``
    (``[n_samples, out_dim*hidden + out_dim]``, torch ``parameters_to_vector`` order)."""
    lib = _lib.load()
    x = _dev_f32(x, "x")
    if x.dim() != 2:
        x = x.reshape(-1, x.shape[-1])
    p, h = x.shape
    sampled_params = _dev_f32(sampled_params, "sampled_params")
    if sampled_params.dim() != 2 or sampled_params.shape[1] != out_dim * h + out_dim:
        raise ValueError("sampled_params must be [n_samples, out_dim*hidden + out_dim]")
    if activation not in _ACTS:
        raise ValueError(f"unknown activation {activation!r}")
    dev = x.device
    out = {"mean": torch.empty(p, out_dim, device=dev), "sigma2": torch.empty(p, out_dim, device=dev)}
    if want_mean2:
        out["mean2"] = torch.empty(p, out_dim, device=dev)
    with _guard(dev):
        _lib.check(lib.ub_laplace_ll_moments(x.data_ptr(), p, h, out_dim, sampled_params.data_ptr(),
                                             sampled_params.shape[0],
                                             _ACTS[activation] | (0 if tensor_cores else 0x100), out["mean"].data_ptr(),
                                             _ptr(out.get("mean2")), out["sigma2"].data_ptr(), _stream()))
    _count(1 if p > 0 else 0)
    return out


def composite_tiles(xys: Tensor, conics: Tensor, opacities: Tensor, colors: Tensor, gaussian_ids: Tensor,
                    tile_bins: Tensor, height: int, width: int, background: Optional[Sequence[float]] = None
                    ) -> Tuple[Tensor, Tensor]:
    """Alpha-composite ``colors [G, CH]`` over pre-binned 16x16 tile lists -> ``([H, W, CH], alpha [H, W, 1])``."""
    lib = _lib.load()
    xys, conics = _dev_f32(xys, "xys"), _dev_f32(conics, "conics")
    opacities = _dev_f32(opacities.reshape(-1), "opacities")
    colors = _dev_f32(colors, "colors")
    g, ch = colors.shape
    if xys.shape != (g, 2) or conics.shape != (g, 3) or opacities.numel() != g:
        raise ValueError("xys [G,2], conics [G,3], opacities [G] must match colors [G,CH]")
    for name, t in (("gaussian_ids", gaussian_ids), ("tile_bins", tile_bins)):
        if t.dtype != torch.int32 or not t.is_cuda:
            raise TypeError(f"{name} must be a CUDA int32 tensor")
    gaussian_ids, tile_bins = gaussian_ids.contiguous(), tile_bins.contiguous()
    tiles = ((width + _lib.UB_TILE - 1) // _lib.UB_TILE) * ((height + _lib.UB_TILE - 1) // _lib.UB_TILE)
    if tile_bins.shape != (tiles, 2):
        raise ValueError(f"tile_bins must be [{tiles}, 2]")
    dev = colors.device
    out = torch.empty(height, width, ch, device=dev)
    alpha = torch.empty(height, width, 1, device=dev)
    bg = (C.c_float * ch)(*([0.0] * ch if background is None else [float(v) for v in background]))
    with _guard(dev):
        _lib.check(lib.ub_composite_tiles(xys.data_ptr(), conics.data_ptr(), opacities.data_ptr(), colors.data_ptr(),
                                          ch, gaussian_ids.data_ptr(), tile_bins.data_ptr(), height, width, bg,
                                          out.data_ptr(), alpha.data_ptr(), _stream()))
    _count(2)  # flat-mean and spread instantiations
    return out, alpha


def composite_tiles_planes(xys: Tensor, conics: Tensor, opacities: Tensor, planes: Sequence[Tensor],
                           gaussian_ids: Tensor, tile_bins: Tensor, height: int, width: int,
                           background: Optional[Sequence[float]] = None, want_max: bool = False
                           ) -> Tuple[List[Tensor], Tensor, Optional[Tensor]]:
    """One compositing pass over several colour planes (``[G, c_p]`` each, <= 8 channels in total) ->
    ``([H, W, c_p] per plane, alpha [H, W, 1], per-channel max keys or None)``."""
    lib = _lib.load()
    xys, conics = _dev_f32(xys, "xys"), _dev_f32(conics, "conics")
    opacities = _dev_f32(opacities.reshape(-1), "opacities")
    g = xys.shape[0]
    pls = [_dev_f32(p.reshape(g, -1), f"planes[{i}]") for i, p in enumerate(planes)]
    chs = [int(p.shape[1]) for p in pls]
    if xys.shape != (g, 2) or conics.shape != (g, 3) or opacities.numel() != g:
        raise ValueError("xys [G,2], conics [G,3], opacities [G] must match the colour planes")
    for name, t in (("gaussian_ids", gaussian_ids), ("tile_bins", tile_bins)):
        if t.dtype != torch.int32 or not t.is_cuda:
            raise TypeError(f"{name} must be a CUDA int32 tensor")
    gaussian_ids, tile_bins = gaussian_ids.contiguous(), tile_bins.contiguous()
    tiles = ((width + _lib.UB_TILE - 1) // _lib.UB_TILE) * ((height + _lib.UB_TILE - 1) // _lib.UB_TILE)
    if tile_bins.shape != (tiles, 2):
        raise ValueError(f"tile_bins must be [{tiles}, 2]")
    dev = xys.device
    total = sum(chs)
    outs = [torch.empty(height, width, c, device=dev) for c in chs]
    alpha = torch.empty(height, width, 1, device=dev)
    keys = torch.empty(total, dtype=torch.int32, device=dev) if want_max else None
    bg = (C.c_float * total)(*([0.0] * total if background is None else [float(v) for v in background]))
    pp = (C.c_void_p * len(pls))(*[p.data_ptr() for p in pls])
    pc = (C.c_int32 * len(pls))(*chs)
    po = (C.c_void_p * len(pls))(*[o.data_ptr() for o in outs])
    with _guard(dev):
        _lib.check(lib.ub_composite_tiles_planes(xys.data_ptr(), conics.data_ptr(), opacities.data_ptr(), pp, pc,
                                                 len(pls), gaussian_ids.data_ptr(), tile_bins.data_ptr(), height,
                                                 width, bg, po, alpha.data_ptr(), _ptr(keys), _stream()))
    _count(2)  # flat-mean and spread instantiations
    return outs, alpha, keys


def composite_tiles_planes_backward(xys: Tensor, conics: Tensor, opacities: Tensor, planes: Sequence[Tensor],
                                    gaussian_ids: Tensor, tile_bins: Tensor, height: int, width: int,
                                    background: Optional[Sequence[float]], v_outs: Sequence[Optional[Tensor]],
                                    v_alpha: Optional[Tensor], want_plane_grads: Optional[Sequence[bool]] = None
                                    ) -> Tuple[Tensor, Tensor, Tensor, List[Optional[Tensor]]]:
    """Gradients of ``composite_tiles_planes`` -> ``(v_xys [G,2], v_conics [G,3], v_opacities [G], [v_plane_p])``."""
    lib = _lib.load()
    xys, conics = _dev_f32(xys, "xys"), _dev_f32(conics, "conics")
    opacities = _dev_f32(opacities.reshape(-1), "opacities")
    g = xys.shape[0]
    dev = xys.device
    pls = [_dev_f32(p.reshape(g, -1), f"planes[{i}]") for i, p in enumerate(planes)]
    chs = [int(p.shape[1]) for p in pls]
    total = sum(chs)
    vos = [None if v is None else _dev_f32(v.reshape(height, width, c), f"v_outs[{i}]")
           for i, (v, c) in enumerate(zip(v_outs, chs))]
    va = None if v_alpha is None else _dev_f32(v_alpha.reshape(height, width), "v_alpha")
    want = [True] * len(pls) if want_plane_grads is None else list(want_plane_grads)
    v_xys = torch.empty(g, 2, device=dev)
    v_conics = torch.empty(g, 3, device=dev)
    v_opac = torch.empty(g, device=dev)
    v_pl = [torch.empty(g, c, device=dev) if w else None for c, w in zip(chs, want)]
    bg = (C.c_float * total)(*([0.0] * total if background is None else [float(v) for v in background]))
    pp = (C.c_void_p * len(pls))(*[p.data_ptr() for p in pls])
    pc = (C.c_int32 * len(pls))(*chs)
    pv = (C.c_void_p * len(pls))(*[_ptr(v) for v in vos])
    pg = (C.c_void_p * len(pls))(*[_ptr(v) for v in v_pl])
    with _guard(dev):
        _lib.check(lib.ub_composite_tiles_planes_backward(
            xys.data_ptr(), conics.data_ptr(), opacities.data_ptr(), pp, pc, len(pls),
            _ptr(gaussian_ids.contiguous()), tile_bins.contiguous().data_ptr(), height, width, bg, pv, _ptr(va), g,
            v_xys.data_ptr(), v_conics.data_ptr(), v_opac.data_ptr(), pg, _stream()))
    _count(2)  # flat-mean and spread instantiations
    return v_xys, v_conics, v_opac, v_pl


def tile_alpha_probe(xys: Tensor, conics: Tensor, opacities: Tensor, gaussian_ids: Tensor, first: int, count: int,
                     tile_x: int, tile_y: int) -> Tuple[Tensor, Tensor]:
    """Diagnostic: ``(sigma, alpha) [count, 16, 16]`` of ``gaussian_ids[first:first + count]`` at the pixel centres of
    tile ``(tile_x, tile_y)``, evaluated by the device function the compositing kernels use (``ub_tile_alpha_probe``)."""
    lib = _lib.load()
    xys, conics = _dev_f32(xys, "xys"), _dev_f32(conics, "conics")
    opacities = _dev_f32(opacities.reshape(-1), "opacities")
    if gaussian_ids.dtype != torch.int32 or not gaussian_ids.is_cuda:
        raise TypeError("gaussian_ids must be a CUDA int32 tensor")
    if first < 0 or count < 0 or first + count > gaussian_ids.numel():
        raise ValueError("probe range outside gaussian_ids")
    dev = xys.device
    sigma = torch.empty(count, _lib.UB_TILE, _lib.UB_TILE, device=dev)
    alpha = torch.empty(count, _lib.UB_TILE, _lib.UB_TILE, device=dev)
    with _guard(dev):
        _lib.check(lib.ub_tile_alpha_probe(xys.data_ptr(), conics.data_ptr(), opacities.data_ptr(),
                                           gaussian_ids.contiguous().data_ptr(), int(first), int(count), int(tile_x),
                                           int(tile_y), sigma.data_ptr(), alpha.data_ptr(), _stream()))
    _count(1 if count > 0 else 0)
    return sigma, alpha


def splat_normalize_(image: Tensor, alpha: Optional[Tensor] = None, max_key: Optional[Tensor] = None,
                     clamp_max_one: bool = False, want_square: bool = False, want_sqrt: bool = False):
    """In place: ``min(image, 1)`` and / or ``alpha > 0 ? image / alpha : max`` (activesplatfacto_model.py:275,319).
    With ``want_square`` / ``want_sqrt`` also returns ``image ** 2`` / ``image.sqrt()`` of the processed image
    (:364, :367) from the same pass: ``image`` alone, or ``(image, square, sqrt)`` with ``None`` for the unwanted."""
    lib = _lib.load()
    ch = int(image.shape[-1])
    n = image.numel() // ch
    sq = torch.empty_like(image) if want_square else None
    rt = torch.empty_like(image) if want_sqrt else None
    with _guard(image.device):
        _lib.check(lib.ub_splat_normalize(image.data_ptr(), ch, _ptr(alpha), n, 1 if clamp_max_one else 0,
                                          1 if alpha is not None else 0, _ptr(max_key), _ptr(sq), _ptr(rt), _stream()))
    _count(1)
    return (image, sq, rt) if (want_square or want_sqrt) else image


def splat_depth_residual(xys: Tensor, depths: Tensor, depth_image: Tensor) -> Tensor:
    """Squared residual of every Gaussian's depth against the rendered depth at its centre pixel -> ``[G, 1]``."""
    lib = _lib.load()
    xys, depths = _dev_f32(xys, "xys"), _dev_f32(depths.reshape(-1), "depths")
    depth_image = _dev_f32(depth_image, "depth_image")
    h, w = int(depth_image.shape[0]), int(depth_image.shape[1])
    out = torch.empty(depths.numel(), 1, device=xys.device)
    with _guard(xys.device):
        _lib.check(lib.ub_splat_depth_residual(xys.data_ptr(), depths.data_ptr(), depth_image.data_ptr(), h, w,
                                               depths.numel(), out.data_ptr(), _stream()))
    _count(2)  # flat-mean and spread instantiations
    return out
