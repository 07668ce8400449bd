#!/usr/bin/env python
"""Benchmark of the uncertainty rendering-and-scoring hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload view|sweep64]

Headline workload (BASELINE.json configs[1], the largest single-GPU configuration): one full 1297 x 840 view,
48 samples per ray, 5 ensemble members -> fused variance compositing of every member (active-nerfacto
outputs, eval chunk 32768) -> per-pixel mean / variance reduce across the members -> AUSE (mae, mse,
rmse) + AUCE + NLL of the view.  One "step" = one such view; at N > 1 every rank takes its own view per
step (weak scaling, views are independent) and the per-view records are all-gathered (NCCL) inside the
timed region.  metric = uncertainty-composited rays/s = members x rays x N / max-over-ranks step time.

The same line carries, under ``extra``, driver-run records of the other BASELINE configs: ``configs0`` (4096-ray
training batch + one 800x800 score: latency-bound), ``configs2`` (K = 10 reduce, tcgen05 Laplace moments, AUCE over
200 views), ``configs3`` (1 M-Gaussian active-splatfacto view) and ``sweep64`` = configs[4] as written: 64 views x 5
members block-sharded over the N ranks (strong scaling), one all_gather of the records, with a SHA-256 of the
gathered records that must be identical at every N.  ``--workload sweep64`` makes that sweep the headline instead.

`--impl reference` times the reference's CPU torch path (the oracle restatement, which tests/test_oracle_pinned.py
pins bit-for-bit to the reference's own code; `/root/reference` itself is not mounted on the GPU box, hence
kind "port") on a bounded sample of the same workload, all host threads.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H, W, S, M = 840, 1297, 48, 5
CHUNK = 1 << 15
BYTES_PER_RAY = 1576            # 48 x 32 B in + 40 B out
METRIC = "uncertainty-composited rays/s"
UNIT = "rays/s"
SWEEP_VIEWS, SWEEP_POOL = 64, 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="view", choices=["view", "sweep64"])
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--members", type=int, default=M)
    ap.add_argument("--views", type=int, default=SWEEP_VIEWS, help="views of the sweep64 workload")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue the step kernel by kernel instead of CUDA-graph replay")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    return ap.parse_args()


def workload_name(h, w, m):
    return (f"configs[1]: {w}x{h} view x {S} samples/ray x {m} active-nerfacto ensemble members -> "
            f"variance compositing (chunk {CHUNK}) -> per-pixel member mean/variance over the 9 image keys "
            f"(branch A: rgb/depth epistemic+aleatoric variance; the [R,S,1] density pass-through is not reduced) "
            f"-> AUSE(mae,mse,rmse)+AUCE+NLL")


def sweep_name(h, w, m, views):
    return (f"configs[4]: {views} views of {w}x{h} x {S} samples/ray x {m} members, block-sharded over the ranks "
            f"-> compositing -> reduce -> AUSE+AUCE+NLL -> one all_gather of the per-view records")


def config_dict(args):
    """The `config` of BOTH arms (identical keys and values, so the driver can tell they ran the same thing)."""
    h, w, m = args.height, args.width, args.members
    if args.workload == "sweep64":
        name = sweep_name(h, w, m, args.views)
    else:
        name = workload_name(h, w, m)
    return {"workload": name, "rays_per_view": h * w, "samples_per_ray": S, "members": m, "rays_per_chunk": CHUNK,
            "l2": f"inputs {m * h * w * 1536 / 1e9:.1f} GB per view >> 126 MB L2, no flush needed"}


# ---------------------------------------------------------------------------------------------------
# clocks: NVML sampled from a thread during the timed region
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device_index: int, period_s: float = float(os.environ.get("UB_CLOCK_PERIOD_MS", "4")) * 1e-3):
        self.samples, self.reasons, self.power = [], set(), []
        self.period = period_s
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover - NVML missing
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        """Start sampling (the thread exists before the timing barrier, so no rank starts late because of it)."""
        if self.nv is not None and self._thread is None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def reset(self):
        self.samples.clear()
        self.reasons.clear()
        self.power.clear()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": getattr(self, "err", "no samples")}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------
# the reference's CPU torch path (oracle restatement), shared by cpu_baseline and --impl reference
def cpu_reference_step(members, gt, h, w):
    from oracle import compositing as oc, metrics as om, reduce as orc
    from uncertainty_nerf_gs_b200.pipeline import PER_SAMPLE_KEYS

    outs = []
    for m in members:
        o = oc.render_in_chunks(oc.active_nerfacto_outputs, CHUNK, m["density"], m["deltas"], m["starts"], m["ends"],
                                m["rgb"], m["beta"])
        outs.append({k: v.view(h, w, -1) for k, v in o.items() if k not in PER_SAMPLE_KEYS})  # same keys as the GPU arm
    red = orc.ensemble_reduce(outs) if len(outs) > 1 else outs[0]
    return om.unc_metrics_rgb(red["rgb"], gt, red["rgb_std"], stable=False)  # the reference's literal sort call


def make_cpu_sample(rows, w, m, seed=0):
    from uncertainty_nerf_gs_b200 import synthetic

    members = [synthetic.ray_samples(rows * w, S, seed=seed * 100 + i) for i in range(m)]
    _, _, gt = synthetic.scoring_image(rows, w, seed=seed)
    return members, gt


def time_cpu_reference(h, w, m, budget_s, steps=1, warmup=0):
    """Time `steps` CPU steps on a sample of whole image rows sized for `budget_s` seconds in total."""
    torch.set_num_threads(os.cpu_count() or 1)
    probe_rows = max(1, 20000 // w)
    members, gt = make_cpu_sample(probe_rows, w, m)
    cpu_reference_step(members, gt, probe_rows, w)
    t0 = time.perf_counter()
    cpu_reference_step(members, gt, probe_rows, w)
    per_px = (time.perf_counter() - t0) / (probe_rows * w)
    rows = int(min(h, max(probe_rows, budget_s / max(1, steps + warmup) / per_px / w)))
    members, gt = make_cpu_sample(rows, w, m)
    for _ in range(warmup):
        cpu_reference_step(members, gt, rows, w)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(members, gt, rows, w)
        times.append(time.perf_counter() - t0)
    rays = rows * w * m
    return {"rows": rows, "rays_per_step": rays, "times": times,
            "sample": f"{rows} of {h} image rows ({rows * w} pixels x {m} members x {S} samples), full pipeline "
                      f"(compositing in {CHUNK}-ray chunks, member reduce, 3x ause + auce + nll)"}


CPU_KIND_NOTE = ("oracle restatement of the reference's torch path (pinned bit-for-bit to the reference's own code by "
                 "tests/test_oracle_pinned.py); the reference package itself is not mounted on the GPU box")


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget = float(os.environ.get("UB_REFERENCE_BUDGET_S", "150"))
    r = time_cpu_reference(args.height, args.width, args.members, budget, steps, warmup)
    t = sum(r["times"]) / len(r["times"])
    value = r["rays_per_step"] / t
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "sweep64" else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "kind_note": CPU_KIND_NOTE,
                         "sample": r["sample"], "cpu_count": os.cpu_count(), "torch": torch.__version__,
                         "numpy": np.__version__},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "images_per_s": 1.0 / (t * (args.height / r["rows"])),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "composite_rays_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


class Ctx:
    """What every section of the run needs."""

    def __init__(self, args, rank, world, local_rank):
        import torch.distributed as dist

        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.dist = dist
        self.dev = torch.device("cuda", local_rank)
        self.peak, self.peak_src = measured_peak()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(self, x: float):
        if self.world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        out = torch.empty(self.world, dtype=torch.float64, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out.tolist()]


def make_view(ctx, h, w, m, slot_seed, gt_seed):
    """Member ray samples of one view (seeded, generated on the device) and a ground truth drawn around its own
    reduced render."""
    from uncertainty_nerf_gs_b200 import pipeline, synthetic
    from uncertainty_nerf_gs_b200.models.outputs import ensemble_reduce

    members = [synthetic.ray_samples(h * w, S, seed=slot_seed * 100 + i, device=ctx.dev) for i in range(m)]
    outs = pipeline.render_members(members, h, w, CHUNK)
    red = ensemble_reduce(outs) if m > 1 else outs[0]
    mean, std = red["rgb"].clone(), red["rgb_std"].clone()
    del outs, red
    return members, mean, std, make_gt(ctx, mean, std, gt_seed)


def make_gt(ctx, mean, std, seed):
    g = torch.Generator(device=ctx.dev).manual_seed(1000 + seed)
    return torch.clamp(mean + std * torch.randn(mean.shape, generator=g, device=ctx.dev), 0.0, 1.0)


def stream_views(ctx, evaluate, view_ids, rows_per_rank, clocks=None, timers_reset=None):
    """THE timed region: barrier, then the stream of views (view i+1 enqueued before view i's record is read
    back), then the path's one exchange.  Returns per-rank timings (device events) and the gathered records."""
    from uncertainty_nerf_gs_b200 import ops, pipeline

    ev0, ev_local, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ctx.barrier()
    if clocks is not None:
        clocks.reset()
    if timers_reset is not None:
        timers_reset()
    launches0 = ops.LAUNCH_COUNT
    torch.cuda.profiler.start()      # no-op unless run under `ncu --profile-from-start off`
    ev0.record()
    pending, local = None, []
    host_enqueue_s = host_finish_s = 0.0
    for vid in view_ids:
        t_a = time.perf_counter()
        nxt = (vid, evaluate(vid))
        t_b = time.perf_counter()
        if pending is not None:
            local.append(pipeline.pack_record(pending[0], pending[1].finish()))
        host_enqueue_s += t_b - t_a
        host_finish_s += time.perf_counter() - t_b
        pending = nxt
    if pending is not None:
        local.append(pipeline.pack_record(pending[0], pending[1].finish()))
    ev_local.record()
    rec = np.stack(local) if local else np.zeros((0, pipeline.RECORD_LEN))
    records = pipeline.gather_records(rec, ctx.dev, rows_per_rank=rows_per_rank) if ctx.world > 1 else rec
    ev1.record()
    ctx.barrier()
    torch.cuda.profiler.stop()
    local_ms, total_ms = ev0.elapsed_time(ev_local), ev0.elapsed_time(ev1)
    n = max(1, len(view_ids))
    return {"elapsed_ms": ctx.max_over_ranks(total_ms), "local_ms_per_rank": ctx.all_ranks(local_ms),
            "gather_ms_per_rank": ctx.all_ranks(total_ms - local_ms), "records": records,
            "launches": ops.LAUNCH_COUNT - launches0,
            "host_ms_per_step": {"enqueue": host_enqueue_s / n * 1e3, "wait_and_tail": host_finish_s / n * 1e3}}


def records_digest(records: np.ndarray) -> str:
    """SHA-256 of the gathered records without their wall-clock fields: identical at every N when the sharded run is
    bit-identical to the single-GPU one."""
    from uncertainty_nerf_gs_b200 import pipeline

    r = np.array(records, dtype=np.float64, copy=True)
    o = 2 * pipeline._CURVES
    for k in pipeline.TIMING_KEYS:
        r[:, o + pipeline.ALL_SCALAR_KEYS.index(k)] = 0.0
    return hashlib.sha256(np.ascontiguousarray(r).tobytes()).hexdigest()


def warm_gather(ctx, rows):
    """The exchange at the size the timed region will use: NCCL sets up its connections / picks its protocol per
    message size lazily, which must not land inside the timed region."""
    from uncertainty_nerf_gs_b200 import pipeline

    if ctx.world > 1:
        fake = np.zeros((rows, pipeline.RECORD_LEN))
        fake[:, -1] = ctx.rank * rows + np.arange(rows)
        for _ in range(2):
            pipeline.gather_records(fake, ctx.dev, rows_per_rank=rows)


def make_evaluator(ctx, h, w, use_graph):
    """`evaluate(members, gt)` -> pending view: CUDA-graph replay of the view's device work (default) or the
    kernel-by-kernel enqueue."""
    from uncertainty_nerf_gs_b200 import pipeline

    timers = []
    if use_graph:
        graphs = pipeline.GraphedViews(h, w, CHUNK)

        def evaluate(members, gt):
            return graphs.launch(members, gt, timers=timers)
    else:
        def evaluate(members, gt):
            return pipeline.evaluate_view_async(members, gt, h, w, CHUNK, timers=timers)
    return evaluate, timers


def run_ours(args, rank, world, local_rank):
    from uncertainty_nerf_gs_b200 import ops, pipeline
    from uncertainty_nerf_gs_b200.build import build_library

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the ub200 path has no CPU fallback")
    ctx = Ctx(args, rank, world, local_rank)
    if rank == 0 or world == 1:
        build_library()
    if world > 1:
        ctx.dist.barrier()
    dev = ctx.dev
    torch.cuda.set_device(dev)
    h, w, m = args.height, args.width, args.members
    R = h * w
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    clocks = ClockSampler(local_rank).start()
    use_graph = not args.no_graph

    sweep = None
    if args.workload == "sweep64" or not args.no_extras:
        sweep = run_sweep(ctx, h, w, m, args.views, use_graph, clocks if args.workload == "sweep64" else None)
        sweep_clocks = clocks.summary()
    if args.workload == "sweep64":
        clocks.stop()
        if rank == 0:
            print(json.dumps(sweep_line(ctx, sweep, sweep_clocks, h, w, m)), flush=True)
        return

    # ---- headline: configs[1], one view per rank per step (seed = view id = rank) ----
    view_id = rank
    members, mean, std, gt = make_view(ctx, h, w, m, view_id, view_id)
    del mean, std
    torch.cuda.synchronize()
    evaluate, timers = make_evaluator(ctx, h, w, use_graph)
    for _ in range(warmup):
        pipeline.pack_record(view_id, evaluate(members, gt).finish())
    warm_gather(ctx, steps)
    run = stream_views(ctx, lambda vid: evaluate(members, gt), [view_id] * steps, steps, clocks,
                       timers_reset=timers.clear)
    clock_summary = clocks.summary()
    clocks.stop()
    elapsed_ms, records = run["elapsed_ms"], run["records"]
    comp_ms = [a.elapsed_time(b) / cnt for a, b, cnt in timers for _ in range(cnt)]   # per compositing launch
    ms_per_step = elapsed_ms / steps
    value = m * R * world / (ms_per_step * 1e-3)

    # ---- the dominant kernel alone: the same compositing calls with nothing else on the device (in the timed
    # region above, the previous view's scoring runs underneath them on a second stream) ----
    solo_timers = []
    for _ in range(3):
        pipeline.render_members(members, h, w, CHUNK, solo_timers)
    torch.cuda.synchronize()
    solo_ms = [a.elapsed_time(b) / cnt for a, b, cnt in solo_timers[1:]]
    comp_solo_ms = sum(solo_ms) / len(solo_ms)
    # the same calls without the deltas stream (ub_composite_rays_args.deltas == NULL: deltas = ends - starts)
    lean_members = [{k: v for k, v in mm.items() if k != "deltas"} for mm in members]
    lean_timers = []
    for _ in range(3):
        pipeline.render_members(lean_members, h, w, CHUNK, lean_timers)
    torch.cuda.synchronize()
    lean_ms = [a.elapsed_time(b) / cnt for a, b, cnt in lean_timers[1:]]
    comp_lean_ms = sum(lean_ms) / len(lean_ms)
    del lean_members

    scoring = time_scoring(ctx, members, gt, h, w, steps)

    # ---- end to end: pinned host inputs -> H2D -> pipeline -> D2H record ----
    e2e = None if args.no_e2e else time_e2e(ctx, members, gt, h, w, m, view_id, steps)

    extras = {}
    if not args.no_extras:
        if sweep is not None:
            extras["sweep64"] = sweep_summary(ctx, sweep, h, w, m)
        if rank == 0 and world == 1:
            del members
            torch.cuda.empty_cache()
            for name, fn in (("configs0", extra_configs0), ("configs2", extra_configs2), ("configs3", extra_configs3)):
                try:
                    extras[name] = fn(ctx)
                except Exception as e:  # an extra must never take the headline down with it
                    extras[name] = {"error": repr(e)}
                torch.cuda.empty_cache()

    if rank != 0:
        return
    agg = pipeline.aggregate_records(records)
    peak, peak_src = ctx.peak, ctx.peak_src
    comp_avg_ms = sum(comp_ms) / len(comp_ms)
    achieved = BYTES_PER_RAY * R / (comp_avg_ms * 1e-3) / 1e9
    cfg = config_dict(args)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "parallelism": {"views_per_step": world, "sharding": f"view-sharded x{world}",
                        "exchange": f"one all_gather of {pipeline.RECORD_LEN * 8} B per-view records at the end of the stream",
                        "step_enqueue": "cuda-graph replay (render graph + score graph per view)" if use_graph
                        else "kernel-by-kernel"},
        "images_per_s": world / (ms_per_step * 1e-3),
        "timing": {"local_ms_per_rank_minmax": [min(run["local_ms_per_rank"]), max(run["local_ms_per_rank"])],
                   "gather_ms_per_rank_minmax": [min(run["gather_ms_per_rank"]), max(run["gather_ms_per_rank"])],
                   "note": "local = barrier -> last record of this rank read back; gather = the all_gather of the "
                           "records incl. waiting for the slowest rank; both inside the timed region"},
        **scoring,
        "roofline": {"kernel": "composite_rays_tma<48,7,14> inside ub_composite_rays_batch (per launch = the batched call of "
                               "M members incl. its one memset and one finalize launch, divided by M)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                     "bytes_per_ray": BYTES_PER_RAY, "rays_per_launch": R, "ms_per_launch": comp_avg_ms,
                     "launches_timed": len(comp_ms), "share_of_step": comp_avg_ms * m / ms_per_step,
                     "traffic": ncu_traffic(),
                     "note": "timed inside the step, where the previous view's scoring kernels share the SMs and HBM "
                             "(second stream); `standalone` = the same call with the device otherwise idle",
                     "standalone": {"ms_per_launch": comp_solo_ms,
                                    "achieved": BYTES_PER_RAY * R / (comp_solo_ms * 1e-3) / 1e9,
                                    "frac": BYTES_PER_RAY * R / (comp_solo_ms * 1e-3) / 1e9 / peak},
                     "derived_deltas": {"ms_per_launch": comp_lean_ms, "bytes_per_ray": BYTES_PER_RAY - 4 * S,
                                        "rays_per_s": R / (comp_lean_ms * 1e-3),
                                        "achieved": (BYTES_PER_RAY - 4 * S) * R / (comp_lean_ms * 1e-3) / 1e9,
                                        "frac": (BYTES_PER_RAY - 4 * S) * R / (comp_lean_ms * 1e-3) / 1e9 / peak,
                                        "note": "standalone, deltas == NULL in ub_composite_rays_args: the kernel takes "
                                                "ends - starts (what RayBundle.get_ray_samples stores as deltas) and reads "
                                                "one stream less -- SURVEY 8(d)'s 1384 B/ray variant; bit-identical "
                                                "outputs for such ray samples.  Not the headline: the reference "
                                                "interface hands deltas over, and so does every other number of this line"}},
        "gpu_launches": run["launches"],
        "host_ms_per_step": run["host_ms_per_step"],
        "clocks": clock_summary,
        "check": {"rgb_ause_rmse": agg["rgb_ause_rmse"], "rgb_nll": agg["rgb_nll"], "views_aggregated": int(records.shape[0])},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if extras:
        line["extra"] = extras
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu_reference(h, w, m, args.cpu_budget_s, steps=1, warmup=0)
        t = r["times"][0]
        line["cpu_baseline"] = {"value": r["rays_per_step"] / t, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "kind_note": CPU_KIND_NOTE, "sample": r["sample"],
                                "cpu_count": os.cpu_count(), "seconds": t, "torch": torch.__version__,
                                "numpy": np.__version__}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def time_scoring(ctx, members, gt, h, w, steps):
    """AUSE + AUCE + NLL images/s on the workload's image size: 8 views per call (select path = the default, and
    through the bit-exact segmented sort), one view per call streamed and synchronous."""
    from uncertainty_nerf_gs_b200 import metrics, pipeline

    pred_img = pipeline.render_members(members[:1], h, w, CHUNK)[0]
    rgb, std = pred_img["rgb"].clone(), pred_img["rgb_std"].clone()
    del pred_img
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s0.record()
        for _ in range(n):
            fn()
        s1.record()
        torch.cuda.synchronize()
        return s0.elapsed_time(s1) / n

    def streamed(args_, n):
        state = {"pend": None}

        def one():
            nxt = metrics.score_rgb_batch_async(*args_)
            if state["pend"] is not None:
                state["pend"].finish()
            state["pend"] = nxt
        ms = timed(one, n)
        state["pend"].finish()
        return ms

    score_steps = max(5, steps)
    one = (rgb, gt, std)
    score_ms = timed(lambda: metrics.score_rgb_batch(*one), score_steps)
    score_stream_ms = streamed(one, score_steps)

    def graph_streamed(args_, n):
        """one view per call, the call being a CUDA-graph replay; two graph slots alternate so that the host tail of
        view i (read-back wait + numpy curves) runs while view i + 1 is on the device"""
        slots = [pipeline.GraphedScore(*args_), pipeline.GraphedScore(*args_)]
        state = {"pend": None, "i": 0}

        def one_():
            nxt = slots[state["i"] & 1].launch()
            state["i"] += 1
            if state["pend"] is not None:
                state["pend"].finish()
            state["pend"] = nxt
        ms = timed(one_, n)
        state["pend"].finish()
        return ms

    score_graph_stream_ms = graph_streamed(one, 4 * score_steps)
    b8 = 8
    eight = tuple(t[None].expand(b8, *t.shape).contiguous() for t in one)
    n8 = max(3, score_steps // 4)
    score_b8_ms = streamed(eight, n8) / b8
    prev = os.environ.get("UB_AUSE_SORT")
    os.environ["UB_AUSE_SORT"] = "1"
    try:
        sort_b8_ms = streamed(eight, n8) / b8
    finally:
        if prev is None:
            os.environ.pop("UB_AUSE_SORT")
        else:
            os.environ["UB_AUSE_SORT"] = prev
    world = ctx.world
    n = h * w
    return {
        "ause_auce_images_per_s": world / (score_b8_ms * 1e-3),
        "ause_auce_images_per_s_sort_path": world / (sort_b8_ms * 1e-3),
        "ause_auce_ms_per_image": score_b8_ms,
        "ause_auce_detail": {
            "mode_of_headline": "8 views of the workload's size per call (segmented launches), calls streamed; AUSE slice "
                                "sums by the sort-free multi-cut select (same element sets as torch.sort(stable=True))",
            "sort_path": "the same call with UB_AUSE_SORT=1: per-image segmented stable radix sort (bit-exact "
                         "permutation) + cut-point prefix sums -- the north-star's kernel (c) as written",
            "eight_views_per_call_ms_per_image": score_b8_ms,
            "eight_views_per_call_ms_per_image_sort_path": sort_b8_ms,
            "streamed_one_view_per_call_ms": score_stream_ms,
            "synchronous_one_view_per_call_ms": score_ms,
            "graph_replay_one_view_per_call_ms": score_graph_stream_ms,
            "images_per_s_one_view_per_call_streamed": world / (score_stream_ms * 1e-3),
            "images_per_s_one_view_per_call_graph_replay": world / (score_graph_stream_ms * 1e-3),
            "one_view_note": "streamed = eager kernel-by-kernel enqueue (host-bound: ~0.14 ms of Python / ctypes + the "
                             "numpy tail per call); graph replay = the same device work as one cudaGraphLaunch, two slots "
                             "alternating (pipeline.GraphedScore)",
            "roofline_select": {"bound": "hbm", "bytes_per_pixel": 92, "model": "prologue 40 B + 3 key reads 36 B + payload reads 16 B",
                                "achieved": 92 * n / (score_b8_ms * 1e-3) / 1e9, "peak": ctx.peak, "unit": "GB/s",
                                "frac": 92 * n / (score_b8_ms * 1e-3) / 1e9 / ctx.peak},
            "roofline_sort": {"bound": "hbm", "bytes_per_pixel": 196, "model": "SURVEY 8(d): prologue 40 B + 4-pass LSD sort/scan model 156 B",
                              "achieved": 196 * n / (sort_b8_ms * 1e-3) / 1e9, "peak": ctx.peak, "unit": "GB/s",
                              "frac": 196 * n / (sort_b8_ms * 1e-3) / 1e9 / ctx.peak}},
    }


def time_e2e(ctx, members, gt, h, w, m, view_id, steps):
    from uncertainty_nerf_gs_b200 import pipeline

    dev, world = ctx.dev, ctx.world
    R = h * w
    host_members = [{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in mm.items()}
                    for mm in members]
    host_gt = torch.empty(gt.shape, dtype=gt.dtype, pin_memory=True).copy_(gt)
    torch.cuda.synchronize()
    ev = pipeline.HostViewEvaluator(m, R, S, h, w, dev)
    e_steps = max(3, min(steps, 10))
    for _ in range(2):
        ev(host_members, host_gt, CHUNK)
    warm_gather(ctx, 1)
    ctx.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e_steps):
        d = ev(host_members, host_gt, CHUNK)
        rec = pipeline.pack_record(view_id, d)[None, :]
        if world > 1:
            pipeline.gather_records(rec, dev, rows_per_rank=1)
    t1.record()
    ctx.barrier()
    e_ms = ctx.max_over_ranks(t0.elapsed_time(t1))
    # the same bytes as a bare pinned->device copy, all ranks at once: what the host side of this box can feed
    slot = ev.slots[0]
    ctx.barrier()
    t0.record()
    for _ in range(2):
        for mh in host_members:
            for k in pipeline.RAY_KEYS:
                slot[k].copy_(mh[k], non_blocking=True)
    t1.record()
    ctx.barrier()
    copy_ms = ctx.max_over_ranks(t0.elapsed_time(t1)) / 2
    copy_gbs = (ev.h2d_bytes - h * w * 12) / (copy_ms * 1e-3) / 1e9
    out = {"value": m * R * world / (e_ms / e_steps * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": ev.h2d_bytes, "d2h_bytes_per_step": ev.d2h_bytes,
           "ms_per_step": e_ms / e_steps, "steps": e_steps,
           "h2d_gbs": ev.h2d_bytes / (e_ms / e_steps * 1e-3) / 1e9,
           "bare_h2d_copy_gbs_per_rank_all_ranks_concurrently": copy_gbs,
           "note": "PCIe / host-memory bound: the pipeline's H2D rate equals the bare pinned->device copy rate measured "
                   "with all ranks copying at once; the per-rank rate falls with N because the ranks share the host's "
                   "memory system (one NUMA node exposed on this box)"}
    # the same call without the deltas stream (HostViewEvaluator(derive_deltas=True)): 12 % fewer bytes over PCIe
    del ev, slot
    torch.cuda.empty_cache()
    ev = pipeline.HostViewEvaluator(m, R, S, h, w, dev, derive_deltas=True)
    for _ in range(2):
        ev(host_members, host_gt, CHUNK)
    ctx.barrier()
    t0.record()
    for _ in range(e_steps):
        d = ev(host_members, host_gt, CHUNK)
        rec = pipeline.pack_record(view_id, d)[None, :]
        if world > 1:
            pipeline.gather_records(rec, dev, rows_per_rank=1)
    t1.record()
    ctx.barrier()
    l_ms = ctx.max_over_ranks(t0.elapsed_time(t1))
    out["derived_deltas"] = {"value": m * R * world / (l_ms / e_steps * 1e-3), "unit": UNIT, "ms_per_step": l_ms / e_steps,
                             "h2d_bytes_per_step": ev.h2d_bytes,
                             "note": "deltas neither copied nor read (ends - starts in the kernel); not the headline"}
    del host_members, host_gt, ev
    return out


# ---------------------------------------------------------------------------------------------------
# configs[4]: the 64-view sweep
def run_sweep(ctx, h, w, m, num_views, use_graph, clocks):
    """64 views x M members, views block-sharded over the ranks (strong scaling).  The ray samples of view v are
    those of pool slot v % 8 (8 distinct member sets = 67 GB resident per rank; 64 distinct sets would be 537 GB),
    its ground truth is drawn per view (seed = view id), so every view has its own record and the gathered records
    depend only on the view ids -- not on how they were sharded."""
    from uncertainty_nerf_gs_b200 import pipeline

    mine = list(pipeline.shard_views(num_views, ctx.rank, ctx.world))
    rows = (num_views + ctx.world - 1) // ctx.world
    pool_n = min(SWEEP_POOL, num_views)
    slots = sorted({v % pool_n for v in mine})
    pool, gts = {}, {}
    for s_ in slots:
        members, mean, std, _ = make_view(ctx, h, w, m, 10_000 + s_, 0)
        pool[s_] = members
        for v in mine:
            if v % pool_n == s_:
                gts[v] = make_gt(ctx, mean, std, v)
        del mean, std
    torch.cuda.synchronize()
    evaluate, timers = make_evaluator(ctx, h, w, use_graph)
    ev = lambda v: evaluate(pool[v % pool_n], gts[v])
    for v in mine[:min(len(mine), 2 * len(slots))]:          # warm-up: every slot's graph captured, >= 3 steps
        ev(v).finish()
    for v in mine[:3]:
        ev(v).finish()
    warm_gather(ctx, rows)
    run = stream_views(ctx, ev, mine, rows, clocks, timers_reset=timers.clear)
    run["num_views"], run["pool"] = num_views, pool_n
    run["comp_ms"] = [a.elapsed_time(b) / cnt for a, b, cnt in timers for _ in range(cnt)]
    del pool, gts
    torch.cuda.empty_cache()
    return run


def sweep_summary(ctx, run, h, w, m):
    from uncertainty_nerf_gs_b200 import pipeline

    nv = run["num_views"]
    agg = pipeline.aggregate_records(run["records"]) if ctx.rank == 0 and len(run["records"]) else {}
    value = m * h * w * nv / (run["elapsed_ms"] * 1e-3)
    return {"workload": sweep_name(h, w, m, nv), "scaling": "strong", "metric": METRIC, "value": value, "unit": UNIT,
            "views": nv, "views_per_rank": (nv + ctx.world - 1) // ctx.world, "n_gpus": ctx.world,
            "elapsed_ms": run["elapsed_ms"], "ms_per_view_per_rank": run["elapsed_ms"] / max(1, (nv + ctx.world - 1) // ctx.world),
            "images_per_s": nv / (run["elapsed_ms"] * 1e-3),
            "local_ms_per_rank_minmax": [min(run["local_ms_per_rank"]), max(run["local_ms_per_rank"])],
            "gather_ms_per_rank_minmax": [min(run["gather_ms_per_rank"]), max(run["gather_ms_per_rank"])],
            "resident_pool": f"{run['pool']} distinct member sets ({run['pool'] * m * h * w * 1536 / 1e9:.0f} GB), view v uses set v % {run['pool']}; "
                             f"ground truth drawn per view",
            "composite_ms_per_launch": (sum(run["comp_ms"]) / len(run["comp_ms"])) if run.get("comp_ms") else None,
            "records_sha256": records_digest(run["records"]) if ctx.rank == 0 else None,
            "records_sha256_note": "identical at N = 1, 2, 4, 8 <=> the sharded run is bit-identical to the single-GPU run",
            "check": {k: agg.get(k) for k in ("rgb_ause_rmse", "rgb_nll", "rgb_auc_abs_error")},
            "views_aggregated": int(len(run["records"]))}


def sweep_line(ctx, run, clock_summary, h, w, m):
    s = sweep_summary(ctx, run, h, w, m)
    comp = run["comp_ms"]
    comp_avg = sum(comp) / len(comp)
    achieved = BYTES_PER_RAY * h * w / (comp_avg * 1e-3) / 1e9
    return {"metric": METRIC, "value": s["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": s["views_per_rank"],
            "warmup": 3, "ms_per_step": s["ms_per_view_per_rank"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(ctx.args),
            "sweep": s, "gpu_launches": run["launches"], "host_ms_per_step": run["host_ms_per_step"],
            "roofline": {"kernel": "composite_rays_tma<48,7,14>", "bound": "hbm", "achieved": achieved, "peak": ctx.peak,
                         "unit": "GB/s", "frac": achieved / ctx.peak, "peak_source": ctx.peak_src,
                         "ms_per_launch": comp_avg, "traffic": ncu_traffic()},
            "clocks": clock_summary}


# ---------------------------------------------------------------------------------------------------
# driver-run records of the other BASELINE configs (rank 0, N = 1)
def _timeit(fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def _best_cpu(fn, n=2):
    fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


def extra_configs0(ctx):
    """configs[0]: 4096 rays x 48 samples (the training batch) + AUSE/AUCE on one 800x800 image.  Latency-bound:
    the kernels take microseconds, so the figures are per-call latencies, kernel-by-kernel and CUDA-graph replay."""
    from oracle import compositing as oc, metrics as om
    from uncertainty_nerf_gs_b200 import metrics, ops, pipeline, synthetic

    dev = ctx.dev
    sets = [synthetic.ray_samples(4096, S, seed=i, device=dev) for i in range(4)]
    call = lambda i: ops.composite_rays(*(sets[i % 4][k] for k in pipeline.RAY_KEYS), rays_per_chunk=CHUNK)
    eager_ms = _timeit(call, 200)
    g = torch.cuda.CUDAGraph()
    call(0)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        keep = call(0)
    graph_ms = _timeit(lambda i: g.replay(), 200)
    p, s_, gt = synthetic.scoring_image(800, 800, seed=0, device=dev)
    score_sync_ms = _timeit(lambda i: metrics.score_rgb_batch(p, gt, s_), 20)
    sg = pipeline.GraphedScore(p, gt, s_)
    score_graph_ms = _timeit(lambda i: sg.launch().finish(), 20)
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_in = {k: v.cpu() for k, v in sets[0].items()}
    cpu_comp = _best_cpu(lambda: oc.active_nerfacto_outputs(**cpu_in))
    pc, sc, gc = p.cpu(), s_.cpu(), gt.cpu()
    cpu_score = _best_cpu(lambda: om.unc_metrics_rgb(pc, gc, sc, stable=False), n=1)
    del keep
    return {"workload": "configs[0]: active-nerfacto variance compositing of 4096 rays x 48 samples + AUSE/AUCE/NLL of one "
                        "800x800 image", "bound": "launch latency (4096 rays = 6.5 MB: ~1 us of HBM time)",
            "composite_call_us": {"kernel_by_kernel": eager_ms * 1e3, "cuda_graph_replay": graph_ms * 1e3},
            "composite_rays_per_s": {"kernel_by_kernel": 4096 / (eager_ms * 1e-3), "cuda_graph_replay": 4096 / (graph_ms * 1e-3)},
            "score_800x800_ms": {"synchronous_call": score_sync_ms, "cuda_graph_replay_incl_readback": score_graph_ms},
            "score_images_per_s": 1e3 / score_graph_ms,
            "cpu": {"composite_4096_rays_ms": cpu_comp * 1e3, "score_800x800_s": cpu_score, "cores": torch.get_num_threads(),
                    "kind": "port"}}


def extra_configs2(ctx):
    """configs[2]: K = 10 MC-dropout reduce at 800x800, last-layer Laplace moments of one 32768-ray chunk
    (tcgen05 rgb head + fp32 density head), AUCE/AUSE over 200 views of 800x800 (20 resident views x 10)."""
    from oracle import laplace as ol, reduce as orc
    from uncertainty_nerf_gs_b200 import metrics, ops, synthetic
    from uncertainty_nerf_gs_b200.models.outputs import mcdropout_reduce

    dev, peak = ctx.dev, ctx.peak
    hh = ww = 800
    n = hh * ww
    K = 10
    passes = synthetic.member_renders(K, hh, ww, seed=0, device=dev)
    red_ms = _timeit(lambda i: mcdropout_reduce(passes), 20)
    c_sum = sum(v.shape[-1] for v in passes[0].values())
    red_bytes = (K * c_sum * 4 + (c_sum + 3) * 4) * n
    P = CHUNK * S
    lap = synthetic.laplace_head(P, 64, 3, 100, seed=0, device=dev)
    theta = lap["mu_q"][None] + lap["eps_draws"] / torch.sqrt(lap["ggn"] + 1.0)[None]
    rgb_ms = _timeit(lambda i: ops.laplace_ll_moments(lap["x"], theta, 3, "sigmoid"), 10)
    th1 = theta[:, :65].contiguous()
    den_ms = _timeit(lambda i: ops.laplace_ll_moments(lap["x"], th1, 1, "exp"), 10)
    B = 20
    imgs = [synthetic.scoring_image(hh, ww, seed=i, device=dev) for i in range(B)]
    pred, std, gt = (torch.stack([im[j] for im in imgs]) for j in range(3))
    del imgs
    state = {"pend": None}

    def one(i):
        nxt = metrics.score_rgb_batch_async(pred, gt, std)
        if state["pend"] is not None:
            state["pend"].finish()
        state["pend"] = nxt
    sc_ms = _timeit(one, 10, warm=2)            # 10 calls x 20 views = the 200-view test set
    state["pend"].finish()
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_passes = [{k: v.cpu() for k, v in p_.items()} for p_ in passes]
    cpu_red = _best_cpu(lambda: orc.mcdropout_reduce(cpu_passes), n=1)
    lx, lt = lap["x"][:32768].cpu(), theta.cpu()
    cpu_lap = _best_cpu(lambda: ol.sample_laplace(lx, lt, 3, torch.sigmoid), n=1)
    flop_rgb, flop_den = 2 * 64 * 3 * 100 * P, 2 * 64 * 1 * 100 * P
    return {"workload": "configs[2]: nerfacto-mcdropout K=10 reduce (800x800, nerfacto image keys) + nerfacto-laplace last-layer "
                        "MC moments of one 32768-ray chunk (1.57 M points, 100 draws) + AUSE/AUCE/NLL over 200 views of 800x800",
            "reduce_k10": {"ms": red_ms, "roofline": {"bound": "hbm", "achieved": red_bytes / (red_ms * 1e-3) / 1e9, "peak": peak,
                                                      "unit": "GB/s", "frac": red_bytes / (red_ms * 1e-3) / 1e9 / peak,
                                                      "bytes_per_pixel": red_bytes // n}},
            "laplace_rgb_head_tcgen05": {"ms": rgb_ms, "points_per_s": P / (rgb_ms * 1e-3),
                                         "roofline": {"bound": "tensor", "achieved": flop_rgb / (rgb_ms * 1e-3) / 1e12,
                                                      "peak": 74.0, "unit": "TFLOP/s", "frac": flop_rgb / (rgb_ms * 1e-3) / 1e12 / 74.0,
                                                      "peak_source": "fp32-FMA roof of the op as the reference computes it (SURVEY 8(d) A3); the "
                                                                     "kernel runs it as 3xTF32 tcgen05 MMAs, so > 1 is possible; MUFU-bound epilogue"}},
            "laplace_density_head_fma": {"ms": den_ms, "points_per_s": P / (den_ms * 1e-3),
                                         "roofline": {"bound": "tensor", "achieved": flop_den / (den_ms * 1e-3) / 1e12, "peak": 74.0,
                                                      "unit": "TFLOP/s", "frac": flop_den / (den_ms * 1e-3) / 1e12 / 74.0,
                                                      "peak_source": "fp32-FMA roof (CUDA cores; the tcgen05 variant of the density head, 2.4x faster, is opt-in: after exp its E[y^2] leaves the 1e-5 contract)"}},
            "score_200_views": {"ms_per_image": sc_ms / B, "images_per_s": B / (sc_ms * 1e-3), "views_per_call": B,
                                "roofline": {"bound": "hbm", "bytes_per_pixel": 92, "achieved": 92 * n * B / (sc_ms * 1e-3) / 1e9,
                                             "peak": peak, "unit": "GB/s", "frac": 92 * n * B / (sc_ms * 1e-3) / 1e9 / peak}},
            "cpu": {"reduce_k10_s": cpu_red, "laplace_rgb_head_points_per_s": 32768 / cpu_lap,
                    "laplace_sample": "32768 points x 100 draws", "cores": torch.get_num_threads(), "kind": "port"}}


def extra_configs3(ctx):
    """configs[3]: active-splatfacto variance-channel alpha compositing, 1 M Gaussians pre-binned into 16x16
    tiles at 1297x840 (binning excluded from the timing, as BASELINE says; its own time reported beside)."""
    from oracle import splat as osp
    from uncertainty_nerf_gs_b200 import binning, ops, synthetic
    from uncertainty_nerf_gs_b200.models.outputs import active_splatfacto_outputs

    dev, peak = ctx.dev, ctx.peak
    G = 1_000_000
    sc = synthetic.splat_scene(G, H, W, seed=0, device=dev)
    ids, bins = binning.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], H, W)
    inter = int(ids.numel())
    bin_ms = _timeit(lambda i: binning.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], H, W), 5, warm=1)
    planes = [sc["rgbs"], sc["betas"], sc["depths"][:, None].contiguous()]
    pass_ms = _timeit(lambda i: ops.composite_tiles_planes(sc["xys"], sc["conics"], sc["opacities"], planes, ids, bins, H, W), 10)
    bgl = [0.1, 0.2, 0.3]

    class _BG:      # the background colour as the model holds it, without a device->host sync per call
        def tolist(self):
            return bgl
    view_ms = _timeit(lambda i: active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"],
                                                          sc["betas"], ids, bins, H, W, _BG()), 10)
    bytes_pass = 48 * inter + 24 * H * W
    small = synthetic.splat_scene(2000, 96, 128, seed=0, mean_scale_px=4.0)
    sid, sbin = osp.bin_gaussians(small["xys"], small["depths"], small["radii"], 96, 128)
    cpu_s = _best_cpu(lambda: osp.rasterize(small["xys"], small["conics"], small["opacities"], small["rgbs"], sid, sbin, 96, 128,
                                            torch.zeros(3)), n=1)
    return {"workload": "configs[3]: active-splatfacto rgb + variance-channel + depth + depth-variance compositing, 1 M Gaussians "
                        f"pre-binned into 16x16 tiles at {W}x{H} ({inter} tile intersections)",
            "fused_pass_rgb_beta_depth": {"ms": pass_ms, "pixels_per_s": H * W / (pass_ms * 1e-3),
                                          "roofline": {"bound": "hbm", "achieved": bytes_pass / (pass_ms * 1e-3) / 1e9, "peak": peak,
                                                       "unit": "GB/s", "frac": bytes_pass / (pass_ms * 1e-3) / 1e9 / peak,
                                                       "bytes_model": "48 B per intersection + 24 B per pixel (SURVEY 8(d) A2)",
                                                       "note": "instruction-bound (per pixel x splat alpha test), not DRAM-bound"}},
            "full_view_ms": view_ms, "views_per_s": 1e3 / view_ms,
            "binning_ms_not_in_view_time": bin_ms,
            "cpu": {"tile_rasterise_128x96_2000_splats_s": cpu_s, "kind": "port",
                    "note": "python loop over tiles x splats; the reference runs gsplat's CUDA kernel here, so this is not a "
                            "reference timing"}}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
