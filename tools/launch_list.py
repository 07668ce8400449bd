"""Per-launch table from an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_*` log: python tools/launch_summary.py file.csv [filter]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
by = collections.OrderedDict()
for d in data:
    by.setdefault((int(d["ID"]), d["Kernel Name"][:64]), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
tot = 0.0
for (i, k), m in by.items():
    if flt and flt not in k:
        continue
    t = m.get("gpu__time_duration.sum", 0) / 1e3
    mb = (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)) / 1e6
    tot += t
    print(f"{i:4d} {k:64s} {t:9.1f} us {mb:9.1f} MB")
print(f"total {tot:.1f} us")
