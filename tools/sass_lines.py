"""Per-source-line warp-instruction counts of one kernel: joins the SASS page of an ncu report (executed counts per
instruction) with the line table of the built library (nvdisasm -g).

    python tools/sass_lines.py gpurun_out/r2_select_full.ncu-rep sel_classify select_cuts [--top 40]
"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _norm(sig: str) -> str:
    """kernel signature without the differences between ncu and c++filt: "(int)48" vs "48", "(bool)1" vs "true", spaces"""
    sig = re.sub(r"\((int|bool|unsigned int|long)\)", "", sig)
    sig = sig.replace("true", "1").replace("false", "0").replace("ub::", "")
    return re.sub(r"\s+", "", sig.split("(")[0] if "<" not in sig else sig[:sig.rindex(">") + 1])


def main():
    rep, kernel, unit = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # first kernel section only
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name")
    name = rows[start][1]
    hdr = rows[start + 1]
    ia, ie, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
    inst = []
    for r in rows[start + 2:]:
        if not r or r[0] == "Kernel Name":
            break
        inst.append((int(r[ia], 16), int(r[ie]), r[isrc].strip()))
    base = inst[0][0]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", unit, str(ROOT / "uncertainty_nerf_gs_b200" / "libub200.so")], cwd=td,
                       capture_output=True)
        # `unit` is a substring match (composite_tiles also extracts composite_tiles_bwd): disassemble every hit
        dis = "\n".join(subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
                        for cubin in sorted(Path(td).glob("*.cubin")))
    # locate the function by its mangled name containing the kernel string
    line_of = {}
    cur_fn, cur_line, want = None, None, False
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", l) or re.match(r"\s*//-+ \.text\.(\S+)", l)
        if m:
            cur_fn = m.group(1)
            dem = subprocess.run(["c++filt", cur_fn.split(",")[0]], capture_output=True, text=True).stdout.strip()
            want = _norm(dem) == _norm(name)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m and want:
            line_of[int(m.group(1), 16)] = cur_line
    per = collections.Counter()
    tot = 0
    for a, n, s in inst:
        per[line_of.get(a - base)] += n
        tot += n
    print(f"# {name}: {tot} warp instructions executed; top source lines")
    src_cache = {}
    for (key, n) in per.most_common(top):
        text = ""
        if key:
            f = next((p for p in (ROOT / "uncertainty_nerf_gs_b200" / "csrc").glob(key[0])), None)
            if f:
                src_cache.setdefault(f, f.read_text().splitlines())
                text = src_cache[f][key[1] - 1].strip()[:110]
        print(f"{100 * n / tot:5.1f}%  {n:11d}  {key}  {text}")


if __name__ == "__main__":
    main()
