"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid): launches, total and mean time.
usage: python tools/launch_summary.py launches.csv > summary.txt"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    ns = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit") == "us":
        ns *= 1e3
    elif r.get("Metric Unit") == "ms":
        ns *= 1e6
    rows.append((r["Kernel Name"].split("(")[0], r["Grid Size"], ns))
groups = OrderedDict()
for k, g, ns in rows:
    n, t = groups.get((k, g), (0, 0.0))
    groups[(k, g)] = (n + 1, t + ns)
total = sum(t for _, t in groups.values())
print("%d launches, %.3f ms of kernel time" % (len(rows), total / 1e6))
for (k, g), (n, t) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print("%4d launches  grid %-18s %9.3f ms total  %8.4f ms/launch  %5.1f %%  %s" % (n, g, t / 1e6, t / n / 1e6, 100 * t / total, k[:110]))
