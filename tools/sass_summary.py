#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS mnemonics in libmfac.so (runs without a GPU):
UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UBLKCP (bulk copy),
HMMA (legacy mma.sync -- must stay 0).   usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = ROOT / "meanflow_audio_codec_b200" / "libmfac.so"
sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "REDG", "LDG", "STG"]
per, cur, i = collections.OrderedDict(), None, -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        cur = names[i] if i < len(names) else m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    if base in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "REDG", "LDG", "STG"):
        per[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# cuobjdump -sass {so.name}: {len(per)} kernels (sm_100a).  Columns: " + " ".join(WANT))
print("# totals: " + ", ".join(f"{k} {tot[k]}" for k in WANT))
short = lambda n: re.sub(r"\(CUtensorMap_st.*|\(mfac::.*|\(const .*|\(float.*", "", n.replace("mfac::", "").replace("(anonymous namespace)::", "").replace("void ", ""))[:96]
for n, c in sorted(per.items(), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["UTMALDG"], kv[0])):
    if c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"] or c["HMMA"]:
        print(" ".join(f"{c[k]:5d}" for k in WANT) + "  " + short(n))
print("# kernels without tcgen05 / TMA instructions (row kernels, MDCT, AdamW): " + str(sum(1 for c in per.values() if not (c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"]))))
