#!/usr/bin/env python
"""Count SASS instructions per CUDA source line for one kernel (nvdisasm -g output), optionally weighting by the
per-instruction execution counts of an `ncu --page source --csv` dump.  Usage: sass_lines.py all_g.txt <mangled> [src.csv]"""
import csv, re, sys, collections
txt = open(sys.argv[1]).read().split('\n')
name = sys.argv[2]
start = next(i for i, l in enumerate(txt) if l.startswith('.text.' + name + ':'))
line = None
rows = []   # (offset, srcline, sass)
for l in txt[start + 1:]:
    if l.startswith('//-----') or l.startswith('.text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
    if m:
        rows.append((int(m.group(1), 16), line, m.group(2)))
weights = None
if len(sys.argv) > 3:
    r = list(csv.reader(open(sys.argv[3])))
    hi = [i for i, x in enumerate(r) if x and x[0] == 'Address'][0]
    hdr = r[hi]; col = {h: i for i, h in enumerate(hdr)}
    data = [x for x in r[hi + 1:] if len(x) == len(hdr)]
    data.sort(key=lambda x: int(x[0], 16)); base = int(data[0][0], 16)
    weights = {int(x[0], 16) - base: (int(x[col['Instructions Executed']]), int(x[col['# Samples']])) for x in data}
agg = collections.Counter(); aggs = collections.Counter(); n = collections.Counter()
for off, ln, sass in rows:
    w, smp = weights.get(off, (0, 0)) if weights else (1, 0)
    agg[ln] += w; aggs[ln] += smp; n[ln] += 1
tot = sum(agg.values()); tots = sum(aggs.values()) or 1
print(f"total {'executed warp-instr' if weights else 'SASS instr'}: {tot}")
src = {}
for ln, v in agg.most_common(45):
    f, k = ln if ln else ('?', 0)
    if f not in src:
        try: src[f] = open('/root/repo/gym_flock_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = []
    text = src[f][k - 1].strip()[:90] if 0 < k <= len(src[f]) else ''
    print(f"{100*v/tot:5.1f}% smp {100*aggs[ln]/tots:5.1f}%  sass={n[ln]:4d}  {f}:{k:4d}  {text}")
