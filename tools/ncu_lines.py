#!/usr/bin/env python
"""Executed warp-instructions and stall samples per CUDA source line of a capture (ncu --import-source on, kernels built with
-lineinfo).   usage: python tools/ncu_lines.py gpurun_out/<name>.ncu-rep [top]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True).stdout
rows = list(csv.reader(io.StringIO(txt.decode("latin-1"))))
cur, agg, tot, tots = None, {}, 0, 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] == "Line No" or r[2] != "-":
        continue
    try:
        ln, ins, smp = int(r[0]), int(r[7]), int(r[6])
    except ValueError:
        continue
    agg[(cur, ln)] = (ins, smp, r[1].strip()[:100]); tot += ins; tots += smp
print(f"warp-instructions {tot}  samples {tots}")
for (f, ln), (ins, smp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f[:24]:24s} {ln:4d} {ins / tot * 100:5.2f}% smp {smp / max(tots, 1) * 100:5.2f}%  {src}")
