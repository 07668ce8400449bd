#!/usr/bin/env python
"""Benchmark of the uncertainty rendering-and-scoring hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], the largest single-GPU configuration): one full 1297 x 840 view,
48 samples per ray, 5 ensemble members -> fused variance compositing of every member (active-nerfacto
outputs, eval chunk 32768) -> per-pixel mean / variance reduce across the members -> AUSE (mae, mse,
rmse) + AUCE + NLL of the view.  One "step" = one such view; at N > 1 every rank takes its own view per
step (weak scaling, views are independent) and the per-view records are all-gathered (NCCL) inside the
timed region.  metric = uncertainty-composited rays/s = members x rays x N / max-over-ranks step time.

`--impl reference` times the reference's CPU torch path (the oracle restatement + the reference-equal
ause/auce, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--members", type=int, default=M)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    return ap.parse_args()


def workload_name(h, w, m):
    return (f"configs[1]: {w}x{h} view x {S} samples/ray x {m} active-nerfacto ensemble members -> "
            f"variance compositing (chunk {CHUNK}) -> per-pixel member mean/variance over the 9 image keys "
            f"(branch A: rgb/depth epistemic+aleatoric variance; the [R,S,1] density pass-through is not reduced) "
            f"-> AUSE(mae,mse,rmse)+AUCE+NLL")


# ---------------------------------------------------------------------------------------------------
# clocks: NVML sampled from a thread during the timed region
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device_index: int, period_s: float = float(os.environ.get("UB_CLOCK_PERIOD_MS", "10")) * 1e-3):
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

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
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
        "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.height, args.width, args.members), "sample": r["sample"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": r["sample"],
                         "cpu_count": os.cpu_count(), "torch": torch.__version__, "numpy": np.__version__},
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


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    from uncertainty_nerf_gs_b200 import ops, pipeline, synthetic
    from uncertainty_nerf_gs_b200.build import build_library

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the ub200 path has no CPU fallback")
    if rank == 0 or world == 1:
        build_library()
    if world > 1:
        dist.barrier()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_bound = pipeline.bind_host_thread_to_gpu(local_rank) if world > 1 else False
    h, w, m = args.height, args.width, args.members
    R = h * w
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    # ---- synthetic view of this rank (seed = view id), generated on the device ----
    view_id = rank
    members = [synthetic.ray_samples(R, S, seed=view_id * 100 + i, device=dev) for i in range(m)]
    outs = pipeline.render_members(members, h, w, CHUNK)
    from uncertainty_nerf_gs_b200.models.outputs import ensemble_reduce

    red = ensemble_reduce(outs) if m > 1 else outs[0]
    g = torch.Generator(device=dev).manual_seed(1000 + view_id)
    gt = torch.clamp(red["rgb"] + red["rgb_std"] * torch.randn(h, w, 3, generator=g, device=dev), 0.0, 1.0)
    del outs, red
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def finish(pending):
        return pipeline.pack_record(view_id, pending.finish())[None, :]

    def gather(local_records):      # the path's one exchange: fixed-size records of all ranks' views, once
        rec = np.concatenate(local_records, axis=0)
        return pipeline.gather_records(rec, dev) if world > 1 else rec

    def step(timers=None):          # synchronous form (warm-up)
        return gather([finish(pipeline.evaluate_view_async(members, gt, h, w, CHUNK, timers=timers))])

    for _ in range(warmup):
        step()
    barrier()
    timers = []
    launches0 = ops.LAUNCH_COUNT
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.profiler.start()  # no-op unless run under `ncu --profile-from-start off`
        ev0.record()
        # stream of views: view i+1 is enqueued before view i's record is read back; the K records of every
        # rank are exchanged by ONE all_gather at the end of the stream, inside the timed region
        pending, local = None, []
        host_enqueue_s = host_finish_s = 0.0
        for _ in range(steps):
            t_a = time.perf_counter()
            nxt = pipeline.evaluate_view_async(members, gt, h, w, CHUNK, timers=timers)
            t_b = time.perf_counter()
            if pending is not None:
                local.append(finish(pending))
            host_enqueue_s += t_b - t_a
            host_finish_s += time.perf_counter() - t_b
            pending = nxt
        local.append(finish(pending))
        records = gather(local)
        ev1.record()
        barrier()
        torch.cuda.profiler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ops.LAUNCH_COUNT - launches0
    comp_ms = [a.elapsed_time(b) / cnt for a, b, cnt in timers for _ in range(cnt)]   # per compositing launch
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / steps
    value = m * R * world / (ms_per_step * 1e-3)

    # ---- scoring-only throughput (AUSE + AUCE images/s), separate timed loop ----
    from uncertainty_nerf_gs_b200 import metrics

    # ---- the dominant kernel alone: the same compositing calls with nothing else on the device (in the timed
    # region above, the previous view's scoring runs underneath them on a second stream) ----
    solo_timers = []
    for _ in range(2):
        pipeline.render_members(members, h, w, CHUNK, solo_timers)
    torch.cuda.synchronize()
    solo_ms = [a.elapsed_time(b) / cnt for a, b, cnt in solo_timers[1:]]
    comp_solo_ms = sum(solo_ms) / len(solo_ms)

    pred_img = pipeline.render_members(members[:1], h, w, CHUNK)[0]
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        metrics.score_rgb_batch(pred_img["rgb"], gt, pred_img["rgb_std"])
    s0.record()
    score_steps = max(5, steps)
    for _ in range(score_steps):
        metrics.score_rgb_batch(pred_img["rgb"], gt, pred_img["rgb_std"])
    s1.record()
    torch.cuda.synchronize()
    score_ms = s0.elapsed_time(s1) / score_steps
    # the same, streamed like the step above (image i+1 enqueued before image i's record is read back) ...
    s0.record()
    pend = None
    for _ in range(score_steps):
        nxt = metrics.score_rgb_batch_async(pred_img["rgb"], gt, pred_img["rgb_std"])
        if pend is not None:
            pend.finish()
        pend = nxt
    pend.finish()
    s1.record()
    torch.cuda.synchronize()
    score_stream_ms = s0.elapsed_time(s1) / score_steps
    # ... and with 8 views per call (one set of segmented launches), streamed the same way
    b8 = 8
    p8, g8, s8 = (t[None].expand(b8, *t.shape).contiguous() for t in (pred_img["rgb"], gt, pred_img["rgb_std"]))
    metrics.score_rgb_batch(p8, g8, s8)
    n8 = max(3, score_steps // 4)
    s0.record()
    pend = None
    for _ in range(n8):
        nxt = metrics.score_rgb_batch_async(p8, g8, s8)
        if pend is not None:
            pend.finish()
        pend = nxt
    pend.finish()
    s1.record()
    torch.cuda.synchronize()
    score_b8_ms = s0.elapsed_time(s1) / n8 / b8
    del p8, g8, s8

    # ---- end to end: pinned host inputs -> H2D -> pipeline -> D2H record ----
    e2e = None
    if not args.no_e2e:
        host_members = [{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in mm.items()}
                        for mm in members]
        host_gt = torch.empty(gt.shape, dtype=gt.dtype, pin_memory=True).copy_(gt)
        torch.cuda.synchronize()
        ev = pipeline.HostViewEvaluator(m, R, S, h, w, dev)
        e_steps = max(3, min(steps, 10))
        for _ in range(2):
            ev(host_members, host_gt, CHUNK)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(e_steps):
            d = ev(host_members, host_gt, CHUNK)
            rec = pipeline.pack_record(view_id, d)[None, :]
            if world > 1:
                pipeline.gather_records(rec, dev)
        t1.record()
        barrier()
        e_ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": m * R * world / (e_ms / e_steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": ev.h2d_bytes, "d2h_bytes_per_step": ev.d2h_bytes,
               "ms_per_step": e_ms / e_steps, "steps": e_steps,
               "h2d_gbs": ev.h2d_bytes / (e_ms / e_steps * 1e-3) / 1e9}
        del host_members, host_gt, ev

    if rank != 0:
        return
    agg = pipeline.aggregate_records(records)
    peak, peak_src = measured_peak()
    comp_avg_ms = sum(comp_ms) / len(comp_ms)
    achieved = BYTES_PER_RAY * R / (comp_avg_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(h, w, m), "rays_per_view": R, "samples_per_ray": S, "members": m,
                   "views_per_step": world, "l2": f"inputs {m * R * 1536 / 1e9:.1f} GB per step >> 126 MB L2, no flush needed",
                   "parallelism": f"view-sharded x{world}, all_gather of {pipeline.RECORD_LEN * 8} B records",
                   "host_threads_bound_to_gpu_numa_node": bool(numa_bound)},
        "images_per_s": world / (ms_per_step * 1e-3),
        "ause_auce_images_per_s": world / (score_b8_ms * 1e-3),
        "ause_auce_ms_per_image": score_b8_ms,
        "ause_auce_detail": {"mode_of_headline": "8 views of the workload's size per call (segmented launches), calls streamed",
                             "eight_views_per_call_ms_per_image": score_b8_ms,
                             "streamed_one_view_per_call_ms": score_stream_ms,
                             "synchronous_one_view_per_call_ms": score_ms,
                             "images_per_s_one_view_per_call_streamed": world / (score_stream_ms * 1e-3)},
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
                                    "frac": BYTES_PER_RAY * R / (comp_solo_ms * 1e-3) / 1e9 / peak}},
        "gpu_launches": launches,
        "host_ms_per_step": {"enqueue": host_enqueue_s / steps * 1e3, "wait_and_tail": host_finish_s / steps * 1e3},
        "clocks": clocks.summary(),
        "check": {"rgb_ause_rmse": agg["rgb_ause_rmse"], "rgb_nll": agg["rgb_nll"], "views_aggregated": int(records.shape[0])},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu_reference(h, w, m, args.cpu_budget_s, steps=1, warmup=0)
        t = r["times"][0]
        line["cpu_baseline"] = {"value": r["rays_per_step"] / t, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "sample": r["sample"], "cpu_count": os.cpu_count(),
                                "seconds": t, "torch": torch.__version__, "numpy": np.__version__}
    print(json.dumps(line), flush=True)


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
