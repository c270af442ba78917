"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.  usage: launch_summary.py launches.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki])          # drop the argument list
    name = re.sub(r"<unnamed>::", "", name)
    a = agg.setdefault(name[:64], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':66s} {'launches':>8s} {'avg_us':>10s} {'total_ms':>9s} {'share':>6s}")
for k, (n, us) in agg.items():
    print(f"{k:66s} {n:8d} {us / n:10.1f} {us / 1e3:9.2f} {100 * us / tot:5.1f}%")
