#!/usr/bin/env python
"""SASS-level execution profile of a kernel from an .ncu-rep: consecutive instructions with the same execution
count are folded into basic-block-like runs.  usage: python tools/ncu_sass.py rep [min_pct=1.0] [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
sel = ["-k", "regex:" + sys.argv[3], "-c", "1"] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = None, []
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
si, ii, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot_i = sum(float(r[ii] or 0) for r in data); tot_s = sum(float(r[ss] or 0) for r in data)
print(f"instructions {tot_i:.0f} samples {tot_s:.0f} sass-lines {len(data)}")
runs = []
for k, r in enumerate(data):
    c = float(r[ii] or 0)
    if runs and runs[-1][0] == c:
        runs[-1][2] = k; runs[-1][3] += float(r[ss] or 0)
    else:
        runs.append([c, k, k, float(r[ss] or 0)])
for c, a, b, smp in runs:
    n = b - a + 1
    pct = c * n / tot_i * 100
    if pct >= min_pct or smp / max(tot_s, 1) * 100 >= min_pct:
        ops = " ".join(data[k][si].split()[0] for k in range(a, min(b + 1, a + 14)))
        print(f"[{a:5d}-{b:5d}] n={n:4d} exec/instr={c:10.0f} inst%={pct:5.1f} smp%={smp/max(tot_s,1)*100:5.1f}  {ops}")
