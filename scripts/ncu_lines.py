#!/usr/bin/env python
"""Executed warp instructions of an ncu report per source line (top N), with lanes active and stall samples.
usage: ncu_lines.py report.ncu-rep [N] [kernel-id]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; fn = None; nfn = 0
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if fn is not None and r[1] != fn: pass
        fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    try:
        ie = int(r[hdr.index("Instructions Executed")] or 0); te = int(r[hdr.index("Thread Instructions Executed")] or 0)
        ss = int(r[hdr.index("# Samples")] or 0)
    except (ValueError, IndexError):
        continue
    a = agg[(cur, int(r[0]))]
    a[0] += ie; a[1] += te; a[2] += ss; a[3] = r[1]
tot = sum(a[0] for a in agg.values()); stot = sum(a[2] for a in agg.values())
print("total warp instructions", tot, "samples", stot)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{f[:24]:24s} {ln:4d} {a[0]/tot:6.2%} lanes {a[1]/max(a[0],1):5.1f} stall {a[2]/max(stot,1):6.2%}  {a[3][:100]}")
