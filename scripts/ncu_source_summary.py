#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output (SASS view): instructions executed and stall samples
per opcode, the dominant stall reasons, and the hottest instructions.
usage: ncu_source_summary.py file.csv [top_n hottest instructions] [top_n opcodes, default all]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
top_ops = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 6
hdr = next(r for r in rows if r and r[0] == "Address")
data = [dict(zip(hdr, r)) for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
def num(d, k):
    try:
        return float(d[k].replace(",", ""))
    except Exception:
        return 0.0
tot_inst = sum(num(d, "Instructions Executed") for d in data)
tot_samp = sum(num(d, "# Samples") for d in data)
print(f"SASS lines {len(data)}  warp-instructions executed {tot_inst:.0f}  samples {tot_samp:.0f}")
byop = collections.defaultdict(lambda: [0.0, 0.0])
for d in data:
    op = d["Source"].split()[0] if not d["Source"].startswith("@") else d["Source"].split()[1]
    op = op.split(".")[0]
    byop[op][0] += num(d, "Instructions Executed")
    byop[op][1] += num(d, "# Samples")
print("\nper opcode: inst share | sample share")
for op, (i, s) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:top_ops]:
    print(f"  {op:12s} {100 * i / tot_inst:6.2f}%  {100 * s / max(tot_samp, 1):6.2f}%")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("\nstall reasons (all samples):")
for h in sorted(stalls, key=lambda h: -sum(num(d, h) for d in data))[:10]:
    print(f"  {h:26s} {100 * sum(num(d, h) for d in data) / max(tot_samp, 1):6.2f}%")
print("\nhottest instructions by samples:")
for d in sorted(data, key=lambda d: -num(d, "# Samples"))[:top]:
    why = max(stalls, key=lambda h: num(d, h))
    print(f"  {d['Address'][-5:]} {100 * num(d, '# Samples') / max(tot_samp, 1):5.2f}% inst {num(d, 'Instructions Executed'):9.0f} {why:22s} {d['Source'][:90]}")
