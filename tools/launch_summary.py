"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
per kernel: launches, total ms, share, average ms, DRAM GB read / written per launch.
    python tools/launch_summary.py launches.csv ["header text"]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
launch = OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[iid], {"name": r[ik]})
    d[r[im]] = float(r[iv].replace(",", ""))
agg = OrderedDict()
for d in launch.values():
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e6
    a[2] += d.get("dram__bytes_read.sum", 0.0) / 1e9
    a[3] += d.get("dram__bytes_write.sum", 0.0) / 1e9
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("# columns: kernel, launches, total ms, share, avg ms, dram read GB/launch, dram write GB/launch")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:60s} {a[0]:4d} {a[1]:10.3f} {100 * a[1] / tot:6.2f}% {a[1] / a[0]:9.3f} {a[2] / a[0]:8.3f} {a[3] / a[0]:8.3f}")
