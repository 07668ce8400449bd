"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list.
Usage: python tools/launch_summary.py file.csv [--per-launch]"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
launches = collections.OrderedDict()
for r in rows[1:]:
    launches.setdefault(r[ii], {"k": r[ki]})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3}
agg = collections.OrderedDict()
for i, e in launches.items():
    t = e["gpu__time_duration.sum"]; us = t[0] * scale.get(t[1], 1e-3)
    by = sum(e[m][0] * scale.get(e[m][1], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") if m in e)
    if "--per-launch" in sys.argv:
        print(f"{i:>4s} {e['k'][:64]:64s} {us:9.1f} us {by / 1e6:9.1f} MB")
    a = agg.setdefault(e["k"][:64], [0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += by
tot = sum(a[1] for a in agg.values())
print(f"# total device time {tot / 1e3:.3f} ms over {len(launches)} launches")
print(f"{'kernel':64s} {'n':>4s} {'avg_us':>9s} {'total_us':>9s} {'share':>6s} {'dram_MB':>9s} {'GB/s':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:64s} {a[0]:4d} {a[1] / a[0]:9.1f} {a[1]:9.1f} {a[1] / tot * 100:5.1f}% {a[2] / 1e6:9.1f} {a[2] / a[1] / 1e3 if a[1] else 0:7.0f}")
