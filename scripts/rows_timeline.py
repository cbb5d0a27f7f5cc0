#!/usr/bin/env python
"""When do the K_B warps of one 64k-record message start and finish?  (DM_ROWS_TIMELINE=1)
Prints, relative to the first warp's start: percentiles of warp start, end-of-rows, exit, and
the per-SM busy window."""
import ctypes as C
import os
import sys

os.environ["DM_ROWS_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import _make_messages, LINES_PER_MSG
from detectmateservice_b200 import _lib
from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import MONITORED_KEYS

msgs = _make_messages(0, n_msgs=6)
det = DeviceDetector(MONITORED_KEYS, max_batch_bytes=len(msgs[0]) + 4096, max_lines=LINES_PER_MSG + 16, table_log2_slots=16)
d = []
for m in msgs:
    t = torch.zeros(len(m) + 64, dtype=torch.uint8, device="cuda")
    t[:len(m)].copy_(torch.frombuffer(bytearray(m), dtype=torch.uint8))
    d.append(t)
st = torch.cuda.Stream()
sp = st.cuda_stream
det.enqueue_device(d[0].data_ptr(), len(msgs[0]), LINES_PER_MSG, 0, 0, 0, sp)
for i in range(12):
    det.enqueue_device(d[1 + i % 5].data_ptr(), len(msgs[0]), 0, 0, 0, 0, sp)
det.sync()
n = C.c_uint32()
_lib.check(det._lib.dm_debug_rows_timeline(det._h, None, 0, C.byref(n)))
buf = np.zeros(4 * n.value, dtype=np.uint64)
_lib.check(det._lib.dm_debug_rows_timeline(det._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size, C.byref(n)))
t = buf.reshape(-1, 4).astype(np.int64)
t = t[t[:, 1] > 0]
t0 = t[:, 1].min()
start, work, exit_ = (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3
pct = [0, 5, 25, 50, 75, 95, 100]
print("warps", len(t), "SMs", len(set(t[:, 0].tolist())))
print("start  us", np.percentile(start, pct).round(2).tolist())
print("rows done us", np.percentile(work, pct).round(2).tolist())
print("exit   us", np.percentile(exit_, pct).round(2).tolist())
sm_first = {}
sm_last = {}
for smid, a, _, c in zip(t[:, 0].tolist(), start.tolist(), work.tolist(), exit_.tolist()):
    sm_first[smid] = min(sm_first.get(smid, 1e9), a)
    sm_last[smid] = max(sm_last.get(smid, 0), c)
f = np.array(list(sm_first.values())); l = np.array([sm_last[k] for k in sm_first])
print("per-SM first start us", np.percentile(f, pct).round(2).tolist())
print("per-SM last exit  us", np.percentile(l, pct).round(2).tolist())
print("warp lifetime us", np.percentile(exit_ - start, pct).round(2).tolist())
