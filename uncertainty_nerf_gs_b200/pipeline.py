"""Per-view evaluation pipeline: M member renders -> ensemble reduce -> AUSE / AUCE / NLL record.

This is the data-parallel unit of the hot path: one view = M x R ray-sample
batches composited, reduced per pixel and scored; views are independent, so ranks take disjoint
blocks of views and only the fixed-size per-view records are all-gathered (``gather_records``), after
which the reference's *ordered* host-side aggregation (eval_uncertainty.py:920-946, 1070-1077) makes the
multi-GPU result identical to the single-GPU one.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import metrics
from .models import outputs as mo

Tensor = torch.Tensor

RAY_KEYS = ("density", "deltas", "starts", "ends", "rgb", "beta")
# Member outputs that are per *sample*, not per pixel.  The reference's ensemble loop averages them too
# (ensemble_pipeline.py:160-162 runs over every key), but the scorer never reads them
# (eval_uncertainty.py:316-322, 427-428) and SURVEY.md section 8(d) config 2 lists only the image keys, so
# the evaluation driver leaves them out of the member reduce: averaging the [H, W, 48] density pass-through
# alone would move 4.6x the bytes of all nine image keys together.  ``models.outputs.ensemble_reduce`` itself
# (the drop-in for the pipeline method) still reduces whatever keys it is given.
PER_SAMPLE_KEYS = ("density",)

# ---- layout of the per-view float64 record that crosses GPUs -------------------------------------------------
# Everything ``get_average_uncertainty_metrics`` accumulates per image (eval_uncertainty.py:856-893, 920-946,
# 1070-1079): the 6 AUSE curves [100] and 5 AUCE curves [99] of the rgb and of the depth modality, and the
# per-image scalars of ``metrics_dict`` in the reference's insertion order (:684-686 psnr / ssim / lpips,
# :714-725 depth_*, :765-776 rgb_*, :948-952 num_rays_per_sec / fps).  Three flags say which groups are present
# (``eval_depth`` / ``eval_rgb`` / image metrics supplied by the model layer); absent groups are zero-filled and
# left out of the aggregate, like the reference leaves them out of ``metrics.json``.
CURVE_KEYS_100 = ("err_mae", "err_mse", "err_rmse", "err_var_mae", "err_var_mse", "err_var_rmse")
CURVE_KEYS_99 = ("coverage_values", "avg_length_values", "coverage_error_values", "abs_coverage_error_values",
                 "neg_coverage_error_values")
IMAGE_KEYS = ("psnr", "ssim", "lpips")
_MODALITY_SCALARS = ("ause_mse", "ause_mae", "ause_rmse", "mse", "rmse", "nll", "avg_var", "auc_abs_error", "auc_length",
                     "auc_neg_error")
DEPTH_SCALAR_KEYS = tuple("depth_" + k for k in _MODALITY_SCALARS)
SCALAR_KEYS = tuple("rgb_" + k for k in _MODALITY_SCALARS)          # the rgb group (name kept from round 1)
TIMING_KEYS = ("num_rays_per_sec", "fps")
ALL_SCALAR_KEYS = IMAGE_KEYS + DEPTH_SCALAR_KEYS + SCALAR_KEYS + TIMING_KEYS      # metrics.json order
DEPTH_CURVE_KEYS_100 = tuple("depth_" + k for k in CURVE_KEYS_100)
DEPTH_CURVE_KEYS_99 = tuple("depth_" + k for k in CURVE_KEYS_99)
_CURVES = 100 * len(CURVE_KEYS_100) + 99 * len(CURVE_KEYS_99)
_FLAGS = ("has_rgb", "has_depth", "has_image_metrics")
RECORD_LEN = 2 * _CURVES + len(ALL_SCALAR_KEYS) + len(_FLAGS) + 1  # + view id (last)


def render_members(members: Sequence[Dict[str, Tensor]], height: int, width: int, rays_per_chunk: int,
                   timers: Optional[List[Tuple[torch.cuda.Event, torch.cuda.Event]]] = None,
                   keep_per_sample: bool = False) -> List[Dict[str, Tensor]]:
    """Composite every member's ray samples (active-nerfacto ``get_outputs`` per eval chunk) and view the
    per-ray outputs as ``[H, W, C]`` like ``get_outputs_for_camera`` does: one batched call for all members
    (one workspace memset, M compositing kernels, one finalize launch).  Per-sample pass-through keys
    (``PER_SAMPLE_KEYS``) are dropped unless ``keep_per_sample``.  ``timers`` receives ``(start, end, M)``."""
    if timers is not None:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
    outs = mo.active_nerfacto_outputs_many(members, rays_per_chunk=rays_per_chunk, image_hw=(height, width))
    if timers is not None:
        t1.record()
        timers.append((t0, t1, len(members)))     # one bracket around the batched call: elapsed / count per launch
    for o in outs:
        if keep_per_sample:
            o["density"] = o["density"].view(height, width, -1)
        else:
            for k in PER_SAMPLE_KEYS:
                o.pop(k, None)
    return outs


def evaluate_view(members: Sequence[Dict[str, Tensor]], rgb_gt: Tensor, height: int, width: int,
                  rays_per_chunk: int = 1 << 15, min_rgb_std_for_nll: float = 3e-2,
                  timers: Optional[list] = None) -> Dict[str, object]:
    """One view, inputs resident on the device: render M members, reduce, score.  Returns the
    reference's per-image entries (``get_unc_metrics_rgb`` dict + ``metrics_dict`` scalars)."""
    outs = render_members(members, height, width, rays_per_chunk, timers)
    red = mo.ensemble_reduce(outs) if len(outs) > 1 else outs[0]
    d = metrics.score_rgb_batch(red["rgb"], rgb_gt, red["rgb_std"], min_rgb_std_for_nll)[0]
    d.update(metrics.per_image_rgb_scalars(d))
    return d


class PendingView:
    """A view whose device work has been enqueued; ``finish()`` returns the reference's per-image entries."""

    def __init__(self, pending_scores):
        self._pending = pending_scores

    def finish(self) -> Dict[str, object]:
        d = self._pending.finish()[0]
        d.update(metrics.per_image_rgb_scalars(d))
        return d


_SCORE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _score_stream(device) -> "torch.cuda.Stream":
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _SCORE_STREAMS.get(idx)
    if st is None:
        st = _SCORE_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return st


def evaluate_view_async(members: Sequence[Dict[str, Tensor]], rgb_gt: Tensor, height: int, width: int,
                        rays_per_chunk: int = 1 << 15, min_rgb_std_for_nll: float = 3e-2,
                        timers: Optional[list] = None, overlap_scoring: bool = True) -> PendingView:
    """``evaluate_view`` without the final host synchronisation: a driver streaming over a test set enqueues
    view i+1 before reading view i's record back, so the device never waits for the numpy tail.

    With ``overlap_scoring`` the scoring of this view is enqueued on a second stream behind an event: its ~20
    small, latency-bound launches (radix passes over one image, cut sums, ...) then run underneath the next
    view's persistent compositing kernels (which leave 55 KB of shared memory and half the registers of every
    SM free) instead of serialising with them.  The reduced images stay referenced by the returned object until
    ``finish()``, so the caching allocator cannot hand their memory to the other stream early."""
    outs = render_members(members, height, width, rays_per_chunk, timers)
    red = mo.ensemble_reduce(outs) if len(outs) > 1 else outs[0]
    rgb, std = red["rgb"], red["rgb_std"]
    if not overlap_scoring:
        return PendingView(metrics.score_rgb_batch_async(rgb, rgb_gt, std, min_rgb_std_for_nll))
    main = torch.cuda.current_stream(rgb.device)
    side = _score_stream(rgb.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        pending = metrics.score_rgb_batch_async(rgb, rgb_gt, std, min_rgb_std_for_nll)
    for t in (rgb, std, rgb_gt):          # allocated on the main stream, read on the side stream: even if the
        t.record_stream(side)             # caller drops the result unfinished, their memory is not reused early
    pending.keep_alive = (rgb, std, rgb_gt)
    return PendingView(pending)


class ViewStream:
    """Streamed evaluation of a test set, the way ``eval_driver`` and ``bench.py`` drive the path.

    ``push(members, rgb_gt)`` composites and reduces one view on the current stream and returns at once.  Every
    ``score_batch`` views, the reduced images are stacked and scored by ONE set of segmented launches on a second
    stream -- underneath the following views' persistent compositing kernels -- and the batch that was enqueued
    before it is read back (its device work finished long ago, so the host does not stall).  ``push`` and
    ``flush`` return the finished views' entries (the reference's per-image dict + ``metrics_dict`` scalars) in
    view order.  Pays off on long streams (a test set): each batch allocates ~100 MB per view of scratch, and the
    last batch is scored with nothing to hide under, so ``bench.py``'s 30-view timed region keeps the per-view
    form (``evaluate_view_async``)."""

    def __init__(self, height: int, width: int, rays_per_chunk: int = 1 << 15, score_batch: int = 8,
                 min_rgb_std_for_nll: float = 3e-2, timers: Optional[list] = None):
        self.h, self.w, self.chunk = height, width, rays_per_chunk
        self.score_batch = max(1, int(score_batch))
        self.min_std = min_rgb_std_for_nll
        self.timers = timers
        self._views: List[Tuple[Tensor, Tensor, Tensor]] = []   # reduced rgb, rgb_std, gt of the open batch
        self._in_flight: List[object] = []                      # PendingScores of enqueued batches, oldest first

    def _enqueue_batch(self) -> None:
        views, self._views = self._views, []
        dev = views[0][0].device
        main = torch.cuda.current_stream(dev)
        side = _score_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if len(views) == 1:
                rgb, std, gt = (t[None] for t in views[0])
            else:
                rgb, std, gt = (torch.stack([v[i] for v in views]) for i in range(3))
            pending = metrics.score_rgb_batch_async(rgb, gt, std, self.min_std)
        for v in views:                              # main-stream allocations read on the side stream
            for t in v:
                t.record_stream(side)
        pending.keep_alive = (views, rgb, std, gt)
        self._in_flight.append(pending)

    @staticmethod
    def _finish(pending) -> List[Dict[str, object]]:
        out = pending.finish()
        for d in out:
            d.update(metrics.per_image_rgb_scalars(d))
        return out

    def push(self, members: Sequence[Dict[str, Tensor]], rgb_gt: Tensor) -> List[Dict[str, object]]:
        outs = render_members(members, self.h, self.w, self.chunk, self.timers)
        red = mo.ensemble_reduce(outs) if len(outs) > 1 else outs[0]
        self._views.append((red["rgb"], red["rgb_std"], rgb_gt))
        done: List[Dict[str, object]] = []
        if len(self._views) >= self.score_batch:
            self._enqueue_batch()
            while len(self._in_flight) > 1:
                done += self._finish(self._in_flight.pop(0))
        return done

    def flush(self) -> List[Dict[str, object]]:
        if self._views:
            self._enqueue_batch()
        done: List[Dict[str, object]] = []
        while self._in_flight:
            done += self._finish(self._in_flight.pop(0))
        return done


class GraphedScore:
    """CUDA-graph replay of the scorer's device work for fixed input tensors (``score_rgb_batch`` of one shape):
    prologue, AUSE slice sums, packing and the device->host copy are one graph launch instead of ~16 kernel
    launches with their Python / ctypes overhead -- what makes the small, latency-bound configurations
    (one 800x800 image per call) fast.  ``launch()`` returns a ``metrics.PendingScores``."""

    def __init__(self, rgb_pred: Tensor, rgb_gt: Tensor, rgb_std: Tensor, min_rgb_std_for_nll: float = 3e-2,
                 pool=None, stream: Optional["torch.cuda.Stream"] = None):
        self.inputs = (rgb_pred, rgb_gt, rgb_std)
        dev = rgb_pred.device
        self.stream = stream
        metrics.score_rgb_batch(rgb_pred, rgb_gt, rgb_std, min_rgb_std_for_nll)   # warm: lazy loads, cached tables
        torch.cuda.synchronize(dev)
        from . import ops

        self.graph = torch.cuda.CUDAGraph()
        kw = {} if pool is None else {"pool": pool}
        count0 = ops.LAUNCH_COUNT
        with torch.cuda.graph(self.graph, **kw):
            self.packed_dev, self.b, self.n, self.c, self.cuts_one = metrics._score_rgb_device(
                rgb_pred, rgb_gt, rgb_std, min_rgb_std_for_nll)
            self.packed_host = torch.empty(self.packed_dev.shape, dtype=self.packed_dev.dtype, pin_memory=True)
            self.packed_host.copy_(self.packed_dev, non_blocking=True)
        self.kernels = ops.LAUNCH_COUNT - count0
        ops._count(-self.kernels)
        self.done = torch.cuda.Event()
        self._last: Optional[metrics.PendingScores] = None

    def launch(self) -> "metrics.PendingScores":
        if self._last is not None:
            self._last.finish()                     # the pinned result buffer is about to be overwritten
        from . import ops

        self.graph.replay()
        ops._count(self.kernels)
        self.done.record()
        self._last = metrics.PendingScores(self.packed_host, self.packed_dev, self.done, self.b, self.n, self.c,
                                           self.cuts_one)
        return self._last


class _GraphSlot:
    """One of the two buffer sets of a member set in ``GraphedViews``: the compositing graph (main stream) and the
    reduce + score + read-back graph (side stream) with their static outputs."""

    def __init__(self, members, gt_shape, h, w, chunk, min_std, pool, side):
        dev = members[0]["density"].device
        from . import ops

        self.gt = torch.empty(gt_shape, device=dev)
        self.comp, self.post = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        count0 = ops.LAUNCH_COUNT
        with torch.cuda.graph(self.comp, pool=pool):
            outs = render_members(members, h, w, chunk)
        self.outs = outs
        with torch.cuda.graph(self.post, pool=pool, stream=side):
            red = mo.ensemble_reduce(outs) if len(outs) > 1 else outs[0]
            self.packed_dev, self.b, self.n, self.c, self.cuts_one = metrics._score_rgb_device(
                red["rgb"], self.gt, red["rgb_std"], min_std)
            self.packed_host = torch.empty(self.packed_dev.shape, dtype=self.packed_dev.dtype, pin_memory=True)
            self.packed_host.copy_(self.packed_dev, non_blocking=True)
        self.red = red
        self.kernels = ops.LAUNCH_COUNT - count0        # kernels of ours inside the two graphs (one replay launches them all)
        ops._count(-self.kernels)                       # capturing launched nothing
        self.comp_done, self.post_done = torch.cuda.Event(), torch.cuda.Event()
        self.pending: Optional[metrics.PendingScores] = None


class GraphedViews:
    """``evaluate_view_async`` as CUDA-graph replays.  The device work of a view with fixed input tensors never
    changes (all segment / cut tables live on the device), so it is captured once per member set: one graph for the
    batched compositing call (main stream) and one for member reduce + scoring + the packed device->host copy (side
    stream, so that it runs underneath the next view's persistent compositing kernels).  Two buffer sets per member
    set alternate, so view i+1 never overwrites what the scoring of view i still reads; all graphs of a buffer
    set index share one memory pool.  Per view the host issues two graph launches and four event operations
    (~0.03 ms) instead of ~25 kernel launches through ctypes (~0.7 ms).  The ground truth is copied into the slot's
    static buffer (13 MB, device to device) so that views sharing ray samples but not ground truth share graphs."""

    def __init__(self, height: int, width: int, rays_per_chunk: int = 1 << 15, min_rgb_std_for_nll: float = 3e-2):
        self.h, self.w, self.chunk, self.min_std = height, width, rays_per_chunk, min_rgb_std_for_nll
        self._slots: Dict[tuple, _GraphSlot] = {}          # (member tensors, buffer set) -> graphs
        self._count = 0
        self._pools = None
        self._side: Optional["torch.cuda.Stream"] = None
        self._last_post: List[Optional["torch.cuda.Event"]] = [None, None]
        self._warm = False

    @staticmethod
    def _key(members) -> tuple:
        return tuple(0 if m.get(k) is None else m[k].data_ptr() for m in members for k in RAY_KEYS)

    def launch(self, members: Sequence[Dict[str, Tensor]], rgb_gt: Tensor, timers: Optional[list] = None) -> PendingView:
        from . import ops

        dev = rgb_gt.device
        turn = self._count & 1                     # consecutive views alternate between the two buffer sets
        self._count += 1
        mkey = self._key(members)
        slot = self._slots.get((mkey, turn))
        if slot is None:
            if self._pools is None:
                self._pools = (torch.cuda.graph_pool_handle(), torch.cuda.graph_pool_handle())
                self._side = _score_stream(dev)
            if not self._warm:                     # eager once: lazy module loading, cached segment / cut tables
                evaluate_view_async(members, rgb_gt, self.h, self.w, self.chunk, self.min_std).finish()
                self._warm = True
            torch.cuda.synchronize(dev)
            for t in (0, 1):                       # both buffer sets at first sight: no capture later, whatever the phase
                self._slots[(mkey, t)] = _GraphSlot(members, rgb_gt.shape, self.h, self.w, self.chunk, self.min_std,
                                                    self._pools[t], self._side)
            torch.cuda.synchronize(dev)
            slot = self._slots[(mkey, turn)]
        if slot.pending is not None:
            slot.pending.finish()                  # its pinned result buffer is about to be overwritten
        main, side = torch.cuda.current_stream(dev), self._side
        if self._last_post[turn] is not None:      # the previous user of this buffer set has finished reading it
            main.wait_event(self._last_post[turn])
        if timers is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(main)
        slot.comp.replay()
        if timers is not None:
            t1.record(main)
            timers.append((t0, t1, len(members)))
        slot.comp_done.record(main)
        with torch.cuda.stream(side):
            side.wait_event(slot.comp_done)
            slot.gt.copy_(rgb_gt, non_blocking=True)
            slot.post.replay()
            slot.post_done.record(side)
        self._last_post[turn] = slot.post_done
        ops._count(slot.kernels)
        rgb_gt.record_stream(side)
        slot.pending = metrics.PendingScores(slot.packed_host, slot.packed_dev, slot.post_done, slot.b, slot.n, slot.c,
                                             slot.cuts_one)
        return PendingView(slot.pending)


class HostViewEvaluator:
    """End-to-end entry: the caller holds one view's member ray samples and ground truth in *pinned host*
    memory; every call copies them to the device on a copy stream (member m+1 uploads while member m
    composites), runs the pipeline and reads the metric record back.  ``derive_deltas``: the members' ``deltas`` are
    neither copied nor read -- the compositor takes ``ends - starts``, which is what ``RayBundle.get_ray_samples``
    stores there (12 % fewer bytes over PCIe)."""

    def __init__(self, num_members: int, num_rays: int, num_samples: int, height: int, width: int, device,
                 derive_deltas: bool = False):
        self.m, self.r, self.s, self.h, self.w = num_members, num_rays, num_samples, height, width
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        shapes = {"density": (num_rays, num_samples, 1), "deltas": (num_rays, num_samples, 1),
                  "starts": (num_rays, num_samples, 1), "ends": (num_rays, num_samples, 1),
                  "rgb": (num_rays, num_samples, 3), "beta": (num_rays, num_samples, 1)}
        if derive_deltas:
            del shapes["deltas"]
        self.keys = tuple(k for k in RAY_KEYS if k in shapes)
        self.slots = [{k: torch.empty(s, device=self.device) for k, s in shapes.items()}
                      for _ in range(num_members)]
        self.gt_dev = torch.empty(height, width, 3, device=self.device)
        self.h2d_bytes = num_members * sum(int(np.prod(s)) * 4 for s in shapes.values()) + height * width * 3 * 4
        self.d2h_bytes = (400 + 5 + 100) * 8  # the packed curve sums / scalar sums / histogram row

    def __call__(self, members_host: Sequence[Dict[str, Tensor]], gt_host: Tensor,
                 rays_per_chunk: int = 1 << 15) -> Dict[str, object]:
        main = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(main)  # the previous call has finished reading the device slots
        events = []
        with torch.cuda.stream(self.copy_stream):
            for mh, slot in zip(members_host, self.slots):
                for k in self.keys:
                    slot[k].copy_(mh[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                events.append(ev)
            self.gt_dev.copy_(gt_host, non_blocking=True)
        outs = []
        for slot, ev in zip(self.slots, events):
            main.wait_event(ev)
            o = mo.active_nerfacto_outputs(slot["density"], slot.get("deltas"), slot["starts"], slot["ends"], slot["rgb"],
                                           slot["beta"], rays_per_chunk=rays_per_chunk, image_hw=(self.h, self.w))
            for k in PER_SAMPLE_KEYS:
                o.pop(k, None)
            outs.append(o)
        red = mo.ensemble_reduce(outs) if len(outs) > 1 else outs[0]
        main.wait_stream(self.copy_stream)
        d = metrics.score_rgb_batch(red["rgb"], self.gt_dev, red["rgb_std"])[0]
        d.update(metrics.per_image_rgb_scalars(d))
        return d


def pack_record(view_id: int, d: Optional[Dict[str, object]], depth: Optional[Dict[str, object]] = None,
                extra: Optional[Dict[str, float]] = None) -> np.ndarray:
    """Fixed-size float64 record of one view.  ``d``: the rgb entries (``get_unc_metrics_rgb`` curves + the
    ``rgb_*`` scalars of ``metrics.per_image_rgb_scalars``) or None (``eval_rgb = False``); ``depth``: the same
    for the depth modality (curves under their plain names + ``depth_*`` scalars, ``metrics.per_image_depth_scalars``);
    ``extra``: ``psnr / ssim / lpips`` from the model layer and ``num_rays_per_sec / fps``."""
    rec = np.zeros(RECORD_LEN, dtype=np.float64)
    o = 0
    for src in (d, depth):
        for k in CURVE_KEYS_100:
            if src is not None:
                rec[o:o + 100] = np.asarray(src[k], dtype=np.float64).reshape(100)
            o += 100
        for k in CURVE_KEYS_99:
            if src is not None:
                rec[o:o + 99] = np.asarray(src[k], dtype=np.float64).reshape(99)
            o += 99
    extra = extra or {}
    for k in ALL_SCALAR_KEYS:
        src = d if k in SCALAR_KEYS else depth if k in DEPTH_SCALAR_KEYS else extra
        if src is not None and k in src:
            rec[o] = float(src[k])
        o += 1
    rec[o:o + 3] = (d is not None, depth is not None, all(k in extra for k in IMAGE_KEYS))
    rec[-1] = float(view_id)
    return rec


def unpack_record(rec: np.ndarray) -> Tuple[int, Dict[str, object]]:
    """Inverse of ``pack_record``: ``(view id, entries)``; rgb curves under their plain names, depth curves
    with a ``depth_`` prefix, only the groups whose flag is set."""
    flags = rec[-1 - len(_FLAGS):-1]
    has_rgb, has_depth, has_img = (bool(v) for v in flags)
    d: Dict[str, object] = {}
    o = 0
    for present, prefix in ((has_rgb, ""), (has_depth, "depth_")):
        for k in CURVE_KEYS_100:
            if present:
                d[prefix + k] = rec[o:o + 100]
            o += 100
        for k in CURVE_KEYS_99:
            if present:
                d[prefix + k] = rec[o:o + 99]
            o += 99
    for k in ALL_SCALAR_KEYS:
        present = has_rgb if k in SCALAR_KEYS else has_depth if k in DEPTH_SCALAR_KEYS else \
            has_img if k in IMAGE_KEYS else True
        if present:
            d[k] = float(rec[o])
        o += 1
    return int(rec[-1]), d


def bind_host_thread_to_gpu(local_rank: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as closest to GPU ``local_rank`` (one process per
    GPU: keeps every rank's launch thread on its GPU's NUMA node instead of migrating across sockets).  Returns
    False when NVML or the affinity call is unavailable -- purely an optimisation."""
    try:
        import os

        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * i + b for i, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


_PINNED: Dict[tuple, Tensor] = {}


def _pinned(shape, dtype) -> Tensor:
    key = (tuple(shape), dtype)
    t = _PINNED.get(key)
    if t is None:
        if len(_PINNED) > 8:
            _PINNED.clear()
        t = _PINNED[key] = torch.empty(shape, dtype=dtype, pin_memory=True)
    return t


def gather_records(local: np.ndarray, device=None, rows_per_rank: Optional[int] = None) -> np.ndarray:
    """All-gather the ``[local_views, RECORD_LEN]`` float64 records of every rank -- the only collective on the
    path (NCCL ``all_gather_into_tensor`` over NVLink when ``device`` is CUDA, gloo on CPU) -- and return them
    ordered by view id.  Ranks may hold different numbers of views (``shard_views`` gives the last ranks short or
    empty blocks when ``num_views % world_size != 0``): every block is padded to a common row count with rows
    marked view id -1, which are stripped after the exchange.  ``rows_per_rank`` (e.g. ``ceil(num_views / world)``)
    fixes that row count up front; without it the per-rank counts are exchanged first."""
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=np.float64).reshape(-1, RECORD_LEN)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        allrec = local
    else:
        world = dist.get_world_size()
        on_gpu = device is not None and torch.device(device).type == "cuda"
        if rows_per_rank is None:
            cnt = torch.tensor([local.shape[0]], dtype=torch.int64)
            cnt = cnt.to(device) if on_gpu else cnt
            counts = torch.empty(world, dtype=torch.int64, device=cnt.device)
            dist.all_gather_into_tensor(counts, cnt)
            rows_per_rank = int(counts.max().item())
        if local.shape[0] > rows_per_rank:
            raise ValueError(f"{local.shape[0]} local records exceed rows_per_rank = {rows_per_rank}")
        rows = max(1, int(rows_per_rank))
        if on_gpu:
            stage = _pinned((rows, RECORD_LEN), torch.float64)
            stage.zero_()
            stage[:, -1] = -1.0
            stage[:local.shape[0]] = torch.from_numpy(local)
            t = stage.to(device, non_blocking=True)
            out = torch.empty((world * rows, RECORD_LEN), dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(out, t)
            host = _pinned((world * rows, RECORD_LEN), torch.float64)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            allrec = host.numpy().copy()
        else:
            t = torch.zeros((rows, RECORD_LEN), dtype=torch.float64)
            t[:, -1] = -1.0
            t[:local.shape[0]] = torch.from_numpy(local)
            out = torch.empty((world * rows, RECORD_LEN), dtype=torch.float64)
            dist.all_gather_into_tensor(out, t)
            allrec = out.numpy()
        allrec = allrec[allrec[:, -1] >= 0]
    order = np.argsort(allrec[:, -1], kind="stable")
    return allrec[order]


def aggregate_records(records: np.ndarray) -> Dict[str, object]:
    """The reference's test-set aggregation (eval_uncertainty.py:920-946, 957-1067, 1070-1077) in view order:
    float64 running sums of the curves / num_images, float32 ``torch.mean`` of the per-image python floats.
    Scalars come back in the reference's ``metrics.json`` order; rows with a negative view id (padding) are ignored."""
    records = records[records[:, -1] >= 0]
    n = records.shape[0]
    curves: Dict[str, np.ndarray] = {}
    scal: Dict[str, list] = {}
    for rec in records:
        _, d = unpack_record(rec)
        for k, v in d.items():
            if isinstance(v, float):
                scal.setdefault(k, []).append(v)
            else:
                curves[k] = curves.get(k, np.zeros(len(v))) + v
    out: Dict[str, object] = {k: v / n for k, v in curves.items()}
    for k in ALL_SCALAR_KEYS:
        if k in scal and len(scal[k]) == n:
            out[k] = float(torch.mean(torch.tensor(scal[k])))
    return out


def shard_views(num_views: int, rank: int, world_size: int) -> range:
    """Block distribution of views over ranks."""
    per = (num_views + world_size - 1) // world_size
    return range(min(num_views, rank * per), min(num_views, (rank + 1) * per))
