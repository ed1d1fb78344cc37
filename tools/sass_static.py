#!/usr/bin/env python
"""Static SASS statistics of one kernel from `nvdisasm -g -c x.cubin`: instructions per source function region
(by "inlined at" chains are ignored; attribution is by the innermost file:line) and per line.
Usage: sass_static.py lines.sass <kernel-substring> [topN]"""
import collections, re, sys
txt = open(sys.argv[1]).read().split('\n')
want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
inside = False; cur = None; per = collections.Counter(); n = 0; ops = collections.Counter()
for l in txt:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: inside = want in m.group(1); continue
    if l.startswith('.section') or re.match(r'\s*\.section', l): 
        if '.text.' in l: inside = want in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if 'inlined at' in l and cur is not None: pass
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);', l)
    if m:
        per[cur] += 1; n += 1
        t = m.group(1).split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += 1
print("instructions:", n)
# bucket by line ranges of bc7_core.cuh
src = {}
try:
    for i, l in enumerate(open('vierkant_b200/csrc/bc7_core.cuh'), 1): src[i] = l.rstrip()
except Exception: pass
for k, v in per.most_common(top):
    t = src.get(k[1], '')[:90] if k and k[0] == 'bc7_core.cuh' else ''
    print(f"{v:6d}  {k[0] if k else None}:{k[1] if k else 0}  {t.strip()}")
print({k: v for k, v in ops.most_common(25)})
