#!/usr/bin/env python
"""Break an `ncu --page source --csv` dump (first kernel instance) into runs of SASS instructions with a
similar execution count -- i.e. loops / phases -- with their share of executed warp-instructions and samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
first = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
first.sort(key=lambda r: int(r[0], 16))
base = int(first[0][0], 16)
I, S, SRC, T = col["Instructions Executed"], col["# Samples"], col["Source"], col["Thread Instructions Executed"]
tot = sum(int(r[I]) for r in first)
tots = sum(int(r[S]) for r in first)
print("total warp instr (one launch):", tot, "SASS lines", len(first), "samples", tots)
runs, cur = [], None
for r in first:
    e, a = int(r[I]), int(r[0], 16) - base
    if cur and abs(e - cur['e']) <= 0.15 * max(e, cur['e'], 1):
        cur['n'] += 1; cur['sum'] += e; cur['smp'] += int(r[S]); cur['thr'] += int(r[T]); cur['end'] = a
    else:
        cur = {'start': a, 'end': a, 'e': e, 'n': 1, 'sum': e, 'smp': int(r[S]), 'thr': int(r[T]), 'src': r[SRC]}
        runs.append(cur)
for c in runs:
    if c['sum'] > tot * thresh:
        print(f"{c['start']:6x}-{c['end']:6x} n={c['n']:4d} exec/instr={c['e']:8d} sum={c['sum']:9d} ({100*c['sum']/tot:4.1f}%) "
              f"smp={c['smp']:5d} ({100*c['smp']/max(tots,1):4.1f}%) thr/warp={c['thr']/max(c['sum'],1):5.1f} {c['src'][:44]}")
