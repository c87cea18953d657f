#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

  launch list :  ncu_summarize.py launches <launches.csv>            (from --metrics gpu__time_duration.sum --csv)
  full capture:  ncu_summarize.py full <prof.ncu-rep>                (reads it with `ncu -i ... --page raw --csv`)

The launch list is cold-cache and serialised: compare SHARES of the step, not absolute times
(/opt/skills/guides/B200_PROFILING.md).
"""
from __future__ import annotations

import collections
import csv
import io
import re
import subprocess
import sys


def short(name: str) -> str:
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("mfac::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\((int|bool)\)", "", name)
    return name[:110]


def launches(path: str) -> None:
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            ns = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") == "us":
                ns *= 1e3
            elif r.get("Metric Unit") == "ms":
                ns *= 1e6
            rows.append((short(r["Kernel Name"]), ns))
    tot = sum(ns for _, ns in rows)
    agg = collections.OrderedDict()
    for k, ns in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    print(f"# {len(rows)} launches, {tot / 1e3:.1f} us total (serialised, cold cache)")
    print(f"{'share':>7} {'n':>5} {'avg_us':>9}  kernel")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * ns / tot:6.2f}% {n:5d} {ns / n / 1e3:9.2f}  {k}")
    fam = collections.Counter()
    for k, ns in rows:
        fam["gemm_tcgen05" if "gemm_tcgen05" in k else ("mdct/imdct" if "mdct" in k else ("adamw" if "adamw" in k else "row kernels"))] += ns
    print("# family shares: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in fam.most_common()))


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def full(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("## " + short(r[idx["Kernel Name"]]))
        for w in WANT:
            if w in idx:
                print(f"   {w:<70} {r[idx[w]]:>16} {units[idx[w]]}")
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            print(f"   {'traffic = dram read + write':<70} {rd + wr:>16.3f} {units[idx['dram__bytes_read.sum']]}")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
