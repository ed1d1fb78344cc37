#!/usr/bin/env python
"""Dynamic profile by code region from `ncu -i X.ncu-rep --page source --csv --print-source sass`:
SASS instructions in address order, cut into regions of (roughly) equal execution count -- i.e. loop nests -- with each
region's share of executed warp instructions and of stall samples, its per-instruction sample rate (1.0 = kernel
average), the top stall reasons and the opcode mix.
Usage: ncu_regions.py sass.csv [min_share_percent]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(r for r in rows if r and r[0] == "Address")
col = {n: k for k, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ins = []
for r in rows:
    if r and r[0].startswith("0x"):
        try:
            st = {n: int(r[col[n]]) for n in stall_cols}
            ins.append((int(r[0], 16), r[1].strip(), int(r[col["Instructions Executed"]]), int(r[col["# Samples"]]), st))
        except ValueError:
            pass
ins.sort()
total = sum(i[2] for i in ins) or 1
total_s = sum(i[3] for i in ins) or 1
entry = ins[0][2] or 1  # the first instruction runs once per warp
regions, cur = [], []
for i in ins:
    if cur:
        mean = sum(x[2] for x in cur) / len(cur)
        if not (0.75 * mean <= i[2] <= 1.33 * mean) and len(cur) >= 4:
            regions.append(cur)
            cur = []
    cur.append(i)
if cur:
    regions.append(cur)
print(f"{len(ins)} SASS instructions, {total:,} executed warp instructions ({total / entry:,.0f} per warp), {total_s:,} samples")
print(f"{'inst%':>6} {'smpl%':>6} {'rate':>5} {'static':>6} {'x/warp':>8}  top stalls | opcodes")
for reg in regions:
    n = sum(x[2] for x in reg)
    if 100.0 * n / total < min_share:
        continue
    s = sum(x[3] for x in reg)
    st = collections.Counter()
    ops = collections.Counter()
    for x in reg:
        for k, v in x[4].items():
            st[k.replace("stall_", "")] += v
        t = x[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += x[2]
    stot = sum(st.values()) or 1
    tops = " ".join(f"{k}:{100 * v // stot}" for k, v in st.most_common(4))
    topo = " ".join(f"{k}:{100 * v // n}" for k, v in ops.most_common(5))
    rate = (s / total_s) / (n / total)
    print(f"{100.0 * n / total:5.1f}% {100.0 * s / total_s:5.1f}% {rate:5.2f} {len(reg):6d} {n / len(reg) / entry:8.1f}  {tops} | {topo}  @{reg[0][0] & 0xFFFFF:05x}")
