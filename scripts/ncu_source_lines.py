#!/usr/bin/env python
"""Per-source-line executed warp instructions of an ncu report, in file order (all launches summed).
usage: ncu_source_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    try:
        ie = int(r[hdr.index("Instructions Executed")] or 0); te = int(r[hdr.index("Thread Instructions Executed")] or 0)
        sm = int(r[hdr.index("# Samples")] or 0)
    except (ValueError, IndexError):
        continue
    k = (cur, int(r[0]))
    a = agg.setdefault(k, [r[1].strip(), 0, 0, 0])
    a[1] += ie; a[2] += te; a[3] += sm
tot = sum(a[1] for a in agg.values()); tots = sum(a[3] for a in agg.values())
print(f"total warp-instr {tot:,} samples {tots:,}")
for (f, ln), (src, ie, te, sm) in sorted(agg.items()):
    if ie / tot * 100 >= minp or sm / max(tots, 1) * 100 >= minp * 2:
        print(f"{f[:14]:14s}:{ln:4d} inst {ie/tot:6.2%} thr {te/max(ie,1):5.1f} smp {sm/max(tots,1):6.2%} | {src[:100]}")
