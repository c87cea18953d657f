#!/usr/bin/env python
"""Top stall sites from `ncu -i rep --page source --csv` output (one kernel).  usage: ncu_hot.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == hdr[0]:
        break
    data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci['# Samples']]) for r in data)
print(rows[0][1][:150] if rows[0] else '', 'total samples', tot, 'instrs', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(data, key=lambda r: -int(r[ci['# Samples']]))[:n]:
    s = {h[6:]: int(r[ci[h]]) for h in stalls if int(r[ci[h]]) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(r[ci['# Samples']].rjust(6), r[ci['Instructions Executed']].rjust(8), r[ci['Source']].strip()[:80].ljust(80), s)
