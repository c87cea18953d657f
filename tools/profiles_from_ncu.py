#!/usr/bin/env python
"""Builds the committed profiles/ summaries from the ncu CSVs of one bench run.
usage: profiles_from_ncu.py <launches.csv> <gemm_metrics.csv> <tag> [per-GPU batch]"""
import collections, csv, io, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from ncu_summarize import short

ROOT = Path(__file__).resolve().parent.parent
launch_csv, gemm_csv, tag = sys.argv[1:4]
batch = sys.argv[4] if len(sys.argv) > 4 else "37888"


def rows_of(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))


# ---- launch list of one step (between the last two MDCT launches)
rows = rows_of(launch_csv)
idx = [i for i, r in enumerate(rows) if "mdct512" in r["Kernel Name"]]
step = rows[idx[-2]:idx[-1]]
agg, fam, tot = collections.OrderedDict(), collections.Counter(), 0.0
for r in step:
    ns = float(r["Metric Value"].replace(",", ""))
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns; tot += ns
    fam["gemm_tcgen05" if "gemm_tcgen05" in k else ("mdct" if "mdct" in k else ("adamw" if "adamw" in k else "row kernels"))] += ns
out = [f"# ncu --metrics gpu__time_duration.sum, ONE training step at per-GPU batch {batch} ({tag})",
       f"# {len(step)} launches, {tot / 1e3:.1f} us serialised (cold cache: compare shares, not absolutes)",
       f"{'share':>7} {'n':>4} {'avg_us':>9}  kernel"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{100 * ns / tot:6.2f}% {n:4d} {ns / n / 1e3:9.2f}  {k[:110]}")
out.append("# family shares: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in fam.most_common()))
(ROOT / "profiles" / f"r01_launches_step_b{batch}.txt").write_text("\n".join(out) + "\n")
keep = ["ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value"]
with open(ROOT / "profiles" / f"r01_launches_step_b{batch}_raw.csv", "w", newline="") as f:
    w = csv.writer(f); w.writerow(keep)
    for r in step:
        w.writerow([r[c] for c in keep])
print("\n".join(out[:12]))

# ---- DRAM traffic / tensor activity of the GEMM launches
rows = rows_of(gemm_csv)
by = collections.OrderedDict()
for r in rows:
    d = by.setdefault(r["ID"], {"k": short(r["Kernel Name"])})
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
    d[r["Metric Name"]] = v
agg = collections.OrderedDict(); tb = tt = 0.0
for d in by.values():
    a = agg.setdefault(d["k"], [0, 0.0, 0.0, 0.0])
    b = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    t = d["gpu__time_duration.sum"]
    a[0] += 1; a[1] += t; a[2] += b; a[3] += d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * t
    tb += b; tt += t
out = [f"# ncu (dram bytes, duration, tensor-pipe activity) for {len(by)} consecutive tcgen05 GEMM launches of a training step, per-GPU batch {batch} ({tag}; one step has 129)",
       f"# total: {tt:.0f} us serialised, {tb / 1e9:.2f} GB DRAM traffic, {tb / len(by) / 1e6:.1f} MB per launch",
       f"{'n':>4} {'avg_us':>8} {'MB/launch':>10} {'GB/s':>7} {'tensor%':>8}  kernel"]
for k, (n, t, b, tp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{n:4d} {t / n:8.1f} {b / n / 1e6:10.1f} {b / t / 1e3:7.0f} {tp / t:8.1f}  {k[:100]}")
(ROOT / "profiles" / "r01_gemm_step_dram_tensor.txt").write_text("\n".join(out) + "\n")
json.dump({"gemm_tcgen05_dram_bytes_per_launch": tb / len(by), "launches": len(by), "total_dram_bytes": tb,
           "source": "profiles/r01_gemm_step_dram_tensor.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum over consecutive GEMM launches of one step, per-GPU batch " + batch + ")"},
          open(ROOT / "profiles" / "ncu_traffic.json", "w"), indent=1)
print("\n".join(out))
