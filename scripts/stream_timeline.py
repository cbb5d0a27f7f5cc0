#!/usr/bin/env python
"""When do the CTAs of one stream-kernel launch start / finish their rows / exit, and how long do
the epilogue's phases take?  (DM_STREAM_TIMELINE=1; development tool, run under gpurun)"""
import ctypes as C
import os
import sys

os.environ["DM_STREAM_TIMELINE"] = "1"
os.environ.setdefault("DM_KERNEL", "stream")
if len(sys.argv) > 2 and sys.argv[2] == "varlen":
    os.environ.setdefault("DM_STREAM_RECHECK", "chain")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import _make_messages, LINES_PER_MSG
from detectmateservice_b200 import _lib
from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import MONITORED_KEYS

overlap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if len(sys.argv) > 2 and sys.argv[2] == "varlen":            # BASELINE config 5 record lengths
    from detectmateservice_b200.synth import AuditSynth, SEED
    g = AuditSynth(SEED + 3)
    msgs = [g.batch_varlen(LINES_PER_MSG, inject=False)[0]] + [g.batch_varlen(LINES_PER_MSG, inject=True)[0] for _ in range(5)]
else:
    msgs = _make_messages(0, n_msgs=6)
det = DeviceDetector(MONITORED_KEYS, max_batch_bytes=max(len(m) for m in msgs) + 4096, max_lines=LINES_PER_MSG + 16, table_log2_slots=16)
det.set_overlap(bool(overlap))
d = []
for m in msgs:
    t = torch.zeros(len(m) + 64, dtype=torch.uint8, device="cuda")
    t[:len(m)].copy_(torch.frombuffer(bytearray(m), dtype=torch.uint8))
    d.append(t)
st = torch.cuda.Stream()
sp = st.cuda_stream
det.enqueue_device(d[0].data_ptr(), len(msgs[0]), LINES_PER_MSG, 0, 0, 0, sp)
for i in range(12):
    det.enqueue_device(d[1 + i % 5].data_ptr(), len(msgs[1 + i % 5]), 0, 0, 0, 0, sp)
det.sync()
n = C.c_uint32()
_lib.check(det._lib.dm_debug_rows_timeline(det._h, None, 0, C.byref(n)))
buf = np.zeros(4 * n.value + 8, dtype=np.uint64)
_lib.check(det._lib.dm_debug_rows_timeline(det._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size, C.byref(n)))
t = buf[:4 * n.value].reshape(-1, 4).astype(np.int64)
epi = buf[4 * n.value:].astype(np.int64)
t0 = t[:, 1].min()
start, rows, exit_ = (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3
pct = [0, 5, 25, 50, 75, 95, 100]
print("overlap", overlap, "CTAs", len(t), "SMs", len(set(t[:, 0].tolist())))
print("start     us", np.percentile(start, pct).round(2).tolist())
print("rows done us", np.percentile(rows, pct).round(2).tolist())
print("exit      us", np.percentile(exit_, pct).round(2).tolist())
print("CTA rows time us", np.percentile(rows - start, pct).round(2).tolist())
print("epilogue stamps us (enter, order wait done, header, zero-fill, alerts, released):", ((epi[:6] - t0) / 1e3).round(2).tolist())
