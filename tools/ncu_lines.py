#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` dump by CUDA source line.

usage: ncu_lines.py dump.csv [top_n]
Prints, per source line (file:line), its share of executed warp instructions, the average number of
active threads per instruction and its share of stall samples."""
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = "?"
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or not r[0].strip().isdigit():
        continue
    try:
        e = int(r[hdr["Instructions Executed"]] or 0)
        t = int(r[hdr["Thread Instructions Executed"]] or 0)
        s = int(r[hdr["# Samples"]] or 0)
    except (ValueError, IndexError):
        continue
    out.append((cur_file, int(r[0]), r[1].strip(), e, t, s))
tot_e = sum(o[3] for o in out) or 1
tot_t = sum(o[4] for o in out)
tot_s = sum(o[5] for o in out) or 1
print(f"total warp-inst {tot_e:.4g}  thread-inst {tot_t:.4g}  avg active threads {tot_t / tot_e:.2f}  samples {tot_s}")
byfile = {}
for f, ln, src, e, t, s in out:
    a = byfile.setdefault(f, [0, 0, 0])
    a[0] += e; a[1] += t; a[2] += s
for f, (e, t, s) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  file {f:32s} warp-inst {100 * e / tot_e:5.1f}%  thr/inst {t / max(e, 1):5.1f}  samples {100 * s / tot_s:5.1f}%")
for f, ln, src, e, t, s in sorted(out, key=lambda o: -o[3])[:topn]:
    print(f"{f[:22]:22s}:{ln:4d} inst {100 * e / tot_e:5.1f}% thr {t / max(e, 1):5.1f} smp {100 * s / tot_s:5.1f}% | {src[:90]}")
