#!/usr/bin/env python
"""Summarise an `ncu -i X.ncu-rep --page source --csv` dump per source line: stall samples, instructions, smem conflicts.
usage: ncu_source_summary.py dump.csv [top]"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors="replace")))
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cur, hdr, tot = None, None, 0.0
STALLS = ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_math", "stall_mio", "stall_wait", "stall_lg", "stall_not_selected",
          "stall_dispatch", "stall_branch_resolving", "stall_no_inst", "stall_selected")
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    extra = len(r) - len(hdr)          # commas inside the source text
    if extra:
        r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
    d = dict(zip(hdr, r))
    try:
        ln = int(r[0])
        s = float(r[4] or 0)
    except ValueError:
        continue
    key = (cur, ln, r[1].strip()[:80])
    a = agg[key]
    a["samples"] += s
    a["inst"] += float(d["Instructions Executed"] or 0)
    a["excess"] += float(d.get("L1 Wavefronts Shared Excessive") or 0)
    a["wave"] += float(d.get("L1 Wavefronts Shared") or 0)
    for k in STALLS:
        a[k] += float(d.get(k) or 0)
    tot += s
ti = sum(v["inst"] for v in agg.values())
tw = sum(v["wave"] for v in agg.values())
te = sum(v["excess"] for v in agg.values())
print(f"total samples {tot:.0f}  warp-instructions {ti:.0f}  smem wavefronts {tw:.0f} (excessive {te:.0f})")
allst = collections.defaultdict(float)
for v in agg.values():
    for k in STALLS:
        allst[k] += v[k]
print("stall mix:", {k: round(100 * v / tot, 1) for k, v in sorted(allst.items(), key=lambda x: -x[1]) if v / tot > 0.005})
for k, v in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:top]:
    st = {a[6:]: round(100 * b / tot, 1) for a, b in v.items() if a.startswith("stall") and b / tot > 0.004}
    print(f"{k[0]}:{k[1]:4d} {100 * v['samples'] / tot:5.1f}% inst {100 * v['inst'] / ti:5.1f}% exc {v['excess'] / max(te, 1) * 100:4.1f}% | {k[2][:64]} | {st}")
