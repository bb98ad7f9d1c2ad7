#!/usr/bin/env python
"""Per-source-line hot spots of a kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_source.py rep [top=30]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
if not hdr:
    print(out[:2000]); sys.exit(1)
si, ii, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
li = hdr.index("Address") if "Address" in hdr else 0
tot_i = sum(float(r[ii] or 0) for r in data)
tot_s = sum(float(r[ss] or 0) for r in data)
print(f"total instructions {tot_i:.0f}  samples {tot_s:.0f}")
data.sort(key=lambda r: -float(r[ss] or 0))
for r in data[:top]:
    print(f"{float(r[ss] or 0)/max(tot_s,1)*100:5.1f}% smp {float(r[ii] or 0)/max(tot_i,1)*100:5.1f}% inst  L{r[li]:>5}: {r[si].strip()[:130]}")
