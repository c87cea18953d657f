#!/usr/bin/env python
"""Builds the committed profiles/ summaries from the ncu CSVs of one bench run.

usage: profiles_from_ncu.py <round-tag e.g. r02> <launches.csv> [<gemm_metrics.csv> <shapes.csv> <csrc_hash.txt>] [per-GPU batch]

* launch list   -> profiles/<round>_launches_step_b<batch>.txt (+ raw csv): every kernel of ONE step with its share
* GEMM metrics  -> profiles/<round>_gemm_step_dram_tensor.txt: per GEMM family the measured DRAM bytes per launch next to
                   the ALGORITHMIC bytes (operands once + the epilogue's streams, from the launch shapes the library logged)
                   and their ratio -- traffic well above 1.0x means re-reads
                -> profiles/ncu_traffic.json with the hash of the CUDA sources the capture ran on (bench.py refuses a
                   capture whose hash is not the build's)
"""
import collections, csv, io, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from ncu_summarize import short

ROOT = Path(__file__).resolve().parent.parent
rnd, launch_csv = sys.argv[1:3]
rest = sys.argv[3:]
gemm_csv = shapes_csv = hash_txt = None
if len(rest) >= 3:
    gemm_csv, shapes_csv, hash_txt = rest[:3]
    rest = rest[3:]
batch = rest[0] if rest else "37888"


def rows_of(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def ns_of(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(r.get("Metric Unit", "ns"), 1)


# ---- launch list of one step (between the last two tokeniser launches)
rows = rows_of(launch_csv)
idx = [i for i, r in enumerate(rows) if "mdct512" in r["Kernel Name"] or "tokenize_prep" in r["Kernel Name"]]
step = rows[idx[-2]:idx[-1]]
agg, fam, tot = collections.OrderedDict(), collections.Counter(), 0.0
for r in step:
    ns = ns_of(r)
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns; tot += ns
    fam["gemm_tcgen05" if "gemm_tcgen05" in k else ("mdct" if "mdct" in k or "tokenize" in k else ("adamw" if "adamw" in k else "row kernels"))] += ns
out = [f"# ncu --metrics gpu__time_duration.sum, ONE training step at per-GPU batch {batch} ({rnd})",
       f"# {len(step)} launches, {tot / 1e3:.1f} us serialised (cold cache: compare shares, not absolutes)",
       f"{'share':>7} {'n':>4} {'avg_us':>9}  kernel"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{100 * ns / tot:6.2f}% {n:4d} {ns / n / 1e3:9.2f}  {k[:110]}")
out.append("# family shares: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in fam.most_common()))
(ROOT / "profiles" / f"{rnd}_launches_step_b{batch}.txt").write_text("\n".join(out) + "\n")
keep = ["ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value"]
with open(ROOT / "profiles" / f"{rnd}_launches_step_b{batch}_raw.csv", "w", newline="") as f:
    w = csv.writer(f); w.writerow(keep)
    for r in step:
        w.writerow([r[c] for c in keep])
print("\n".join(out[:14]))
if not gemm_csv:
    sys.exit(0)

# ---- algorithmic bytes per GEMM launch from the shapes the library logged (family 0 = tcgen05 GEMM)
# epilogue streams in bytes per output element, by functor label (csrc/epilogues.cuh, imf_kernels.cuh)
EPI_BYTES = {"bias_gelu": 4, "mul_dgelu": 4, "linear_bf16": 2, "linear_f32": 4, "block_out": 12, "block_out_tangent": 14,
             "grad_store": 4, "store_f32": 4, "affine_residual": 10}
shapes = []
for line in open(shapes_csv):
    f = line.strip().split(",")
    if len(f) >= 7 and f[0] == "0":
        M, N, K = int(f[2]), int(f[3]), int(f[4])
        shapes.append((f[1], M, N, K, 2.0 * (M * K + K * N) + EPI_BYTES.get(f[1], 4) * M * N))

# ---- DRAM traffic / tensor activity of the GEMM launches
rows = rows_of(gemm_csv)
by = collections.OrderedDict()
for r in rows:
    d = by.setdefault(r["ID"], {"k": short(r["Kernel Name"])})
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}[u]
    d[r["Metric Name"]] = v
launches = list(by.values())
aligned = len(launches) == len(shapes)
agg = collections.OrderedDict(); tb = tt = ta = 0.0
for i, d in enumerate(launches):
    label = shapes[i][0] if aligned else "?"
    a = agg.setdefault((d["k"], label), [0, 0.0, 0.0, 0.0, 0.0])
    b = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    t = d["gpu__time_duration.sum"]
    alg = shapes[i][4] if aligned else 0.0
    a[0] += 1; a[1] += t; a[2] += b; a[3] += d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * t; a[4] += alg
    tb += b; tt += t; ta += alg
out = [f"# ncu (dram bytes, duration, tensor-pipe activity) for {len(by)} consecutive tcgen05 GEMM launches = one training step, per-GPU batch {batch} ({rnd})",
       f"# total: {tt:.0f} us serialised, {tb / 1e9:.2f} GB DRAM traffic, {tb / len(by) / 1e6:.1f} MB per launch; algorithmic {ta / 1e9:.2f} GB "
       f"(operands once + epilogue streams) -> traffic ratio {tb / ta if ta else float('nan'):.2f}x" + ("" if aligned else "  [shape log not aligned: no algorithmic bytes]"),
       f"{'n':>4} {'avg_us':>8} {'MB/launch':>10} {'alg MB':>8} {'ratio':>6} {'GB/s':>7} {'tensor%':>8}  kernel / epilogue"]
for (k, label), (n, t, b, tp, alg) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{n:4d} {t / n:8.1f} {b / n / 1e6:10.1f} {alg / n / 1e6:8.1f} {b / alg if alg else float('nan'):6.2f} {b / t / 1e3:7.0f} {tp / t:8.1f}  {k[:80]} / {label}")
(ROOT / "profiles" / f"{rnd}_gemm_step_dram_tensor.txt").write_text("\n".join(out) + "\n")
json.dump({"gemm_tcgen05_dram_bytes_per_launch": tb / len(by), "launches": len(by), "total_dram_bytes": tb,
           "algorithmic_bytes": ta if aligned else None, "traffic_ratio": (tb / ta) if (aligned and ta) else None,
           "csrc_hash": open(hash_txt).read().strip(),
           "source": f"profiles/{rnd}_gemm_step_dram_tensor.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum over the consecutive GEMM launches of one step, per-GPU batch {batch})"},
          open(ROOT / "profiles" / "ncu_traffic.json", "w"), indent=1)
print("\n".join(out))
