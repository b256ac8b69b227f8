"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python profiles/launch_table.py gpurun_out/launches.csv
"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], newline="")) if r and r[0].strip('"').isdigit()]
hdr = None
for r in csv.reader(open(sys.argv[1], newline="")):
    if r and r[0] == "ID":
        hdr = r
        break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[ik].split("(")[0].replace("se3::", "").replace("void ", "")
    if "cub::" in name:
        name = "cub::" + name.split("::")[-1].split("<")[0]
    v = float(r[iv].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r[iu]]
    agg[name][0] += 1
    agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f%% | %.1f |" % (k, n, t, 100 * t / tot, t / n))
print("| **total** | %d | %.1f | | |" % (sum(v[0] for v in agg.values()), tot))
