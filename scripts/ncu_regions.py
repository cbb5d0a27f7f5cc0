#!/usr/bin/env python
"""Executed warp instructions of an ncu report summed over source regions.
usage: ncu_regions.py report.ncu-rep file:lo-hi:name ..."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
regions = []
for spec in sys.argv[2:]:
    f, rng, name = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
agg = collections.Counter(); thr = collections.Counter(); tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    try:
        ie = int(r[hdr.index("Instructions Executed")] or 0); te = int(r[hdr.index("Thread Instructions Executed")] or 0)
    except (ValueError, IndexError):
        continue
    ln = int(r[0]); tot += ie
    name = "other:" + cur
    for f, lo, hi, n in regions:
        if cur.startswith(f) and lo <= ln <= hi: name = n; break
    agg[name] += ie; thr[name] += te
for n, v in agg.most_common():
    print(f"{n:34s} {v/tot:6.1%}  thr/inst {thr[n]/max(v,1):5.1f}")
