#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep captured with --import-source on.
Each SASS instruction is counted once (the source view repeats it for every frame of its inline stack).
usage: python tools/ncu_lines.py rep kernel-regex [min_pct=0.7]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur_file = None; cur = None; by = {}; text = {}
nk = 0
for r in rows:
    if len(r) == 2 and r[0] == "Function Name":
        nk += 1
    if nk > 1 and len(r) == 2 and r[0] == "Kernel Name":
        break
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); ss = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr):
        if r[0] != "": cur = (cur_file, int(r[0])); text[cur] = r[1].strip()
        else:
            try: ad = int(r[2], 16)
            except ValueError: continue
            e = by.setdefault(ad, [float(r[ii] or 0), float(r[ss] or 0), r[3].strip(), []]); e[3].append(cur)
agg = {}
for ad, (ic, sc, txt, lines) in by.items():
    key = lines[-1]          # innermost frame listed last
    a = agg.setdefault(key, [0.0, 0.0, 0]); a[0] += ic; a[1] += sc; a[2] += 1
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print(f"instructions {ti:.0f}  samples {ts:.0f}  sass {len(by)}")
for k, a in sorted(agg.items()):
    if a[0] / ti * 100 >= min_pct or a[1] / ts * 100 >= min_pct:
        print(f"{a[1]/ts*100:5.1f}%smp {a[0]/ti*100:5.1f}%inst n={a[2]:3d} {k[0]}:{k[1]}: {text.get(k, '')[:110]}")
