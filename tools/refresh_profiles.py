"""Turn the raw captures of tools/refresh_profiles.sh (gpurun_out/<tag>_*) into the tracked files of profiles/.
Usage: python tools/refresh_profiles.py r1"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout


for name in (f"{tag}_bench_n1.json", f"{tag}_bench_reference.json", f"{tag}_launches.csv", f"{tag}_perf_kernels.jsonl",
             f"{tag}_score16_launches.csv"):
    if os.path.exists(os.path.join(src, name)):
        shutil.copy(os.path.join(src, name), os.path.join(dst, name))
hdr = (f"# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline` (timed region only,\n"
       f"# --profile-from-start off, --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none); raw: {tag}_launches.csv\n"
       f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  2 steps = 2 views x 5 members.\n")
open(os.path.join(dst, f"{tag}_launches_summary.txt"), "w").write(
    hdr + run([sys.executable, "tools/launch_summary.py", f"gpurun_out/{tag}_launches.csv"]))
open(os.path.join(dst, f"{tag}_score16_launches_summary.txt"), "w").write(
    "# ncu launch list of one batched scoring call, 16 views of 800x800 (tools/profile_score.py)\n" +
    run([sys.executable, "tools/launch_summary.py", f"gpurun_out/{tag}_score16_launches.csv", "--per-launch"]))

# ---- ncu --set full summary ----
rep = os.path.join(src, f"{tag}_full.ncu-rep")
if os.path.exists(rep):
    raw = run(["ncu", "-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    lines = [f"# ncu --set full --clock-control none --import-source on, kernels of one bench step ({tag}); units as ncu reports them\n"]
    seen, traffic = {}, None
    for r in rows[2:]:
        k = r[h.index("Kernel Name")]
        if seen.get(k, 0) >= 1:
            continue
        seen[k] = 1
        lines.append(f"\n== {k}\n")
        for w in want:
            if w in h:
                lines.append(f"   {w:86s} {r[h.index(w)]:>16s} {rows[1][h.index(w)]}\n")
        if "composite_rays_tma" in k:
            unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(r[h.index("dram__bytes_read.sum")]) * unit[rows[1][h.index("dram__bytes_read.sum")]]
            wr = float(r[h.index("dram__bytes_write.sum")]) * unit[rows[1][h.index("dram__bytes_write.sum")]]
            tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3}[rows[1][h.index("gpu__time_duration.sum")]]
            traffic = {"kernel": k, "rays_per_launch": 1089480, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": 1089480 * 1576,
                       "duration_us": float(r[h.index("gpu__time_duration.sum")]) * tu,
                       "source": f"ncu --set full, gpurun_out/{tag}_full.ncu-rep"}
    open(os.path.join(dst, f"{tag}_ncu_summary.txt"), "w").writelines(lines)
    if traffic:
        json.dump(traffic, open(os.path.join(dst, "composite_rays_traffic.json"), "w"), indent=1)
# ---- every kernel of one batched scoring call (select path), full metric set ----
rep = os.path.join(src, f"{tag}_select_full.ncu-rep")
if os.path.exists(rep):
    raw = run(["ncu", "-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__inst_executed_op_shared_atom.sum", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    lines = [f"# ncu --set full --clock-control none --import-source on: every kernel of one batched scoring call, 16 views of\n"
             f"# 800x800 through the select path (tools/profile_score.py, {tag}); units as ncu reports them\n"]
    for r in rows[2:]:
        lines.append(f"\n== {r[h.index('Kernel Name')]}\n")
        for w in want:
            if w in h:
                lines.append(f"   {w:86s} {r[h.index(w)]:>16s} {rows[1][h.index(w)]}\n")
    for kern, unit in (("sel_classify", "select_cuts"), ("score_prologue_kernel", "score_prologue")):
        lines.append(f"\n# warp instructions per source line, {kern} (tools/sass_lines.py)\n")
        lines.append(run([sys.executable, "tools/sass_lines.py", rep, kern, unit, "--top", "24"]))
    open(os.path.join(dst, f"{tag}_select_ncu_summary.txt"), "w").writelines(lines)
open(os.path.join(dst, f"{tag}_sass_summary.txt"), "w").write(run(["bash", "tools/sass_summary.sh"]))
for name in (f"{tag}_bench_n2.json", f"{tag}_bench_n4.json", f"{tag}_bench_n8.json", f"{tag}_diag_overlap.txt"):
    if os.path.exists(os.path.join(src, name)):
        shutil.copy(os.path.join(src, name), os.path.join(dst, name))
print("profiles/ refreshed from", src)
