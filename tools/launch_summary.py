#!/usr/bin/env python
"""Per-kernel summary of one training step out of an ncu gpu__time_duration launch list.
usage: launch_summary.py <launches.csv> [top]"""
import collections, csv, io, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from ncu_summarize import short

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
idx = [i for i, r in enumerate(rows) if "mdct512" in r["Kernel Name"]]
step = rows[idx[-2]:idx[-1]]
agg, tot = collections.OrderedDict(), 0.0
for r in step:
    ns = float(r["Metric Value"].replace(",", ""))
    k = short(r["Kernel Name"])[:100] + " g" + r["Grid Size"].replace(" ", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns; tot += ns
print(len(step), "launches", round(tot / 1e3, 1), "us")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 100]:
    print(f"{100 * ns / tot:6.2f}% {n:4d} {ns / n / 1e3:9.2f}  {k}")
