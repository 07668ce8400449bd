"""Per-launch time, instruction count and DRAM bytes of an ncu --csv launch list (tools/score_iter.sh):
    python tools/launch_insts.py gpurun_out/c4_score16.csv [...]"""
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    d, order = {}, []
    for r in rows[1:]:
        k = (r[0], r[ik][:44])
        if k not in d:
            d[k] = {}
            order.append(k)
        d[k][r[im]] = float(r[iv].replace(",", ""))
    print(path)
    tot = 0.0
    for k in order:
        m = d[k]
        t = m["gpu__time_duration.sum"]
        t = t / 1000 if t > 1000 else t
        tot += t
        mb = (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0))
        print(f"  {k[1]:46s} {t:8.1f} us  inst {m.get('smsp__inst_executed.sum', 0) / 1e6:8.2f} M")
    print(f"  total {tot:.1f} us")
