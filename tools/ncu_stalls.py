#!/usr/bin/env python
"""Per-source-line stall-reason samples from `ncu --page source --csv --print-source sass,cuda`.
Usage: ncu_stalls.py src.csv [reason=stall_no_inst] [topN]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
reason = sys.argv[2] if len(sys.argv) > 2 else "stall_no_inst"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
per = collections.Counter(); tot_all = collections.Counter(); src = {}
cols = None; cur = None; cur_file = None
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No": cols = {n: k for k, n in enumerate(r)}
    elif r and r[0] not in ("Function Name",):
        if r[0] != "":
            cur = (cur_file, int(r[0])); src[cur] = r[1].strip()
        elif cols and len(r) > cols[reason] and r[3].strip():
            try: v = int(r[cols[reason]] or 0); s = int(r[cols["# Samples"]] or 0)
            except ValueError: continue
            per[cur] += v; tot_all[cur] += s
T = sum(per.values()); S = sum(tot_all.values())
print(f"{reason}: {T} of {S} samples ({100*T/max(S,1):.1f}%)")
for k, v in per.most_common(top):
    print(f"{100*v/max(T,1):5.1f}%  ({v:6d}/{tot_all[k]:6d} line samples)  {k[0]}:{k[1]:<5} {src.get(k,'')[:100]}")
