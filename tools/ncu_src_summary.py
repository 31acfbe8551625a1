#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall samples and executed instructions by code region
(regions are given as SASS address ranges or found by source-line markers) and the top stalled instructions."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
# first row is the kernel name; find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x") or (len(r) == len(hdr) and r[0].isdigit())]
col = {h: i for i, h in enumerate(hdr)}
S = col["# Samples"]; I = col["Instructions Executed"]; SRC = col["Source"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s = sum(int(r[S] or 0) for r in data); tot_i = sum(int(r[I] or 0) for r in data)
print(f"instructions(SASS lines)={len(data)} total samples={tot_s} total warp-instr executed={tot_i}")
agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print(f"top {top} by samples:")
for r in sorted(data, key=lambda r: -int(r[S] or 0))[:top]:
    st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"  {r[0]:>8s} smp={int(r[S] or 0):6d} exec={int(r[I] or 0):9d} {r[SRC][:70]:70s} {st}")
# coarse regions: cumulative by address order in chunks
print("cumulative by address decile:")
n = len(data)
for d in range(10):
    chunk = data[d * n // 10:(d + 1) * n // 10]
    print(f"  {chunk[0][0]}..{chunk[-1][0]} samples={sum(int(r[S] or 0) for r in chunk):7d} exec={sum(int(r[I] or 0) for r in chunk):10d}")
