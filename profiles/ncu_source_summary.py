#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` : instructions executed and stall samples
per source line (top N) and per SASS opcode.  Usage: ncu_source_summary.py src.csv [topN]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
per_line = collections.Counter()
samples_line = collections.Counter()
per_op = collections.Counter()
samples_op = collections.Counter()
src_text = {}
cur_file, cur_line, total, total_s = None, None, 0, 0
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        cols = {name: k for k, name in enumerate(r)}
        ci, cs = cols["Instructions Executed"], cols["# Samples"]
    elif r and r[0] not in ("Function Name",):
        if r[0] != "":
            cur_line = (cur_file, int(r[0]))
            src_text[cur_line] = r[1].strip()
        elif len(r) > 3 and r[3].strip():
            op = r[3].split()[0] if not r[3].strip().startswith("@") else r[3].split()[1]
            try:
                n, s = int(r[ci]), int(r[cs])
            except ValueError:
                n, s = 0, 0
            per_line[cur_line] += n
            samples_line[cur_line] += s
            per_op[op.split(".")[0]] += n
            samples_op[op.split(".")[0]] += s
            total += n
            total_s += s
    i += 1
print(f"total warp-instructions {total:,}  samples {total_s:,}")
print("\n== top source lines by instructions executed ==")
for k, n in per_line.most_common(top):
    print(f"{100*n/total:5.1f}% inst {100*samples_line[k]/max(total_s,1):5.1f}% smpl  {k[0]}:{k[1]:<5} {src_text.get(k,'')[:110]}")
print("\n== SASS opcodes ==")
for k, n in per_op.most_common(30):
    print(f"{100*n/total:5.1f}% inst {100*samples_op[k]/max(total_s,1):5.1f}% smpl  {k}")
