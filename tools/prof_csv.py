#!/usr/bin/env python
"""Aggregate the per-launch CSV written by libmfac when MFAC_PROFILE_CSV is set (family,label,M,N,K,us,work)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1]))]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
agg = collections.OrderedDict(); tot = 0.0
for r in rows:
    k = tuple(r[:5]); a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += float(r[5]); a[2] += float(r[6]); tot += float(r[5])
print(f"total {tot/steps:.1f} us/step over {len(rows)} launches ({steps:g} steps)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    us = a[1] / a[0]
    rate = f"{a[2]/a[0]/us/1e6:7.0f} TF/s" if k[0] == '0' else f"{a[2]/a[0]/us/1e3:7.0f} GB/s"
    print(f"{a[1]/steps:9.1f} us/step  n={a[0]/steps:5.1f} avg={us:7.1f} us {rate}  fam={k[0]} {k[1]:<18} M={k[2]} N={k[3]} K={k[4]}")
