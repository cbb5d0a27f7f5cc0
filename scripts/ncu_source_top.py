#!/usr/bin/env python
"""Top source lines of an ncu report by executed instructions / stall samples.
usage: ncu_source_top.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0].isdigit():
        d = dict(zip(hdr, r))
        # hdr has duplicate "Source": index manually
        try:
            ie = int(r[hdr.index("Instructions Executed")] or 0)
            te = int(r[hdr.index("Thread Instructions Executed")] or 0)
            sm = int(r[hdr.index("# Samples")] or 0)
        except (ValueError, IndexError):
            continue
        lines.append((cur_file, int(r[0]), r[1].strip(), ie, te, sm))
tot_i = sum(l[3] for l in lines); tot_s = sum(l[5] for l in lines)
print(f"total warp-instr {tot_i:,}  samples {tot_s:,}")
print("--- by instructions executed")
for f, ln, src, ie, te, sm in sorted(lines, key=lambda l: -l[3])[:N]:
    print(f"{ie/tot_i:6.1%} {ie:>10,} thr/inst {te/max(ie,1):5.1f} smp {sm/max(tot_s,1):5.1%}  {f}:{ln}  {src[:90]}")
print("--- by stall samples")
for f, ln, src, ie, te, sm in sorted(lines, key=lambda l: -l[5])[:N//2]:
    print(f"{sm/max(tot_s,1):6.1%} {sm:>8,}  inst {ie/tot_i:5.1%}  {f}:{ln}  {src[:90]}")
