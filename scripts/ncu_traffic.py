#!/usr/bin/env python
"""profiles/r02_stream_traffic.json from an `ncu --set full` report of dm_k_stream: DRAM bytes per launch, tied to
the kernel sources by bench._csrc_sha().   usage: ncu_traffic.py report.ncu-rep out.json"""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _csrc_sha
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
launches = []
for r in rows[2:]:
    if "dm_k_stream" not in r[col("Kernel Name")]:
        continue
    rd = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
    wr = to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
    launches.append({"kernel": r[col("Kernel Name")], "dram_read": rd, "dram_write": wr,
                     "duration_us": float(r[col("gpu__time_duration.sum")]),
                     "inst_executed": float(r[col("smsp__inst_executed.sum")]),
                     "registers": int(float(r[col("launch__registers_per_thread")])), "grid": int(float(r[col("launch__grid_size")]))})
res = {"csrc_sha": _csrc_sha(), "report": os.path.basename(rep), "launches": launches,
       "dram_bytes_per_launch": sum(l["dram_read"] + l["dram_write"] for l in launches) / max(len(launches), 1),
       "note": "ncu replays every launch with flushed caches and one launch at a time: durations are cold and serialised"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res)[:600])
