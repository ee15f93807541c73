#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table).
usage: launch_summary.py launches.csv [title]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "ID")
agg = collections.defaultdict(list)
for r in rows[rows.index(hdr) + 1:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(d["Metric Unit"], v)
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    agg[name[:90]].append(v)
tot = sum(sum(v) for v in agg.values())
if len(sys.argv) > 2:
    print(f"# {sys.argv[2]}")
print("| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {100 * sum(v) / tot:.1f}% |")
