#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total,
max and share.  usage: ncu_summary.py launches.csv [--last N] [--seq]"""
import csv, sys, collections, re
path = sys.argv[1]
last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else None
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    rows.append((name, v * scale))
if last:
    rows = rows[-last:]
if "--seq" in sys.argv:
    for i, (n, ms) in enumerate(rows):
        print("%4d %-40s %9.4f" % (i, n, ms))
    sys.exit(0)
agg = collections.OrderedDict()
for n, ms in rows:
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | max ms | share |\n|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.3f | %.1f %% |" % (n, a[0], a[1], a[2], 100 * a[1] / tot))
print("| total | %d | %.3f | | |" % (len(rows), tot))
