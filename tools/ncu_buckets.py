#!/usr/bin/env python
"""Bucket executed instructions / samples of an ncu source CSV by the device function that contains each source line
of bc7_core.cuh (function = nearest preceding 'VKT_FN ... name(' definition).  Inlined helpers (fadd, prmt ...) are
attributed to themselves.  Usage: ncu_buckets.py src.csv"""
import collections, csv, re, sys
src = open('vierkant_b200/csrc/bc7_core.cuh').read().split('\n')
fn_at = {}
cur = 'top'
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:VKT_FN|VKT_NOINLINE)\s+.*?(\w+)\(', l)
    if m: cur = m.group(1)
    fn_at[i] = cur
rows = list(csv.reader(open(sys.argv[1])))
inst = collections.Counter(); smp = collections.Counter()
cols = None; key = None; cur_file = None
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No": cols = {n: k for k, n in enumerate(r)}
    elif r and r[0] not in ("Function Name",):
        if r[0] != "":
            key = fn_at.get(int(r[0]), '?') if cur_file == 'bc7_core.cuh' else cur_file
        elif cols and len(r) > 8 and r[3].strip():
            try: n = int(r[cols["Instructions Executed"]] or 0); s = int(r[cols["# Samples"]] or 0)
            except ValueError: continue
            inst[key] += n; smp[key] += s
T = sum(inst.values()); S = sum(smp.values())
for k, v in inst.most_common(40):
    print(f"{100*v/T:5.1f}% inst {100*smp[k]/S:5.1f}% smpl  {k}")
