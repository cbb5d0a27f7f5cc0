#!/usr/bin/env python
"""Per-step host timestamps of the pipelined end-to-end loop (collect(slot) then submit(slot))."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import _make_messages, LINES_PER_MSG
from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import MONITORED_KEYS

msgs = _make_messages(0, n_msgs=8)
det = DeviceDetector(MONITORED_KEYS, max_batch_bytes=len(msgs[0]) + 4096, max_lines=LINES_PER_MSG + 16, table_log2_slots=16)
from detectmateservice_b200.numa import bound_to_gpu_node, gpu_numa_cpus
print("gpu-local cpus:", len(gpu_numa_cpus(0) or []), "bind:", os.environ.get("BIND", "1"))
import contextlib
ctx = bound_to_gpu_node(0) if os.environ.get("BIND", "1") == "1" else contextlib.nullcontext()
pin = []
with ctx:
    for m in msgs:
        t = torch.empty(len(m), dtype=torch.uint8, pin_memory=True)
        t.copy_(torch.frombuffer(bytearray(m), dtype=torch.uint8))
        pin.append(t)
    det.process_lines(pin[0].numpy(), LINES_PER_MSG)
    det.submit(pin[1].numpy(), 0, 0); det.collect(0)
N = 60
ts_collect, ts_submit = [], []
for rep in range(2):
    ts_collect, ts_submit = [], []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        if i >= 2:
            det.collect(i & 1)
            ts_collect.append(time.perf_counter() - t0)
        det.submit(pin[1 + i % 7].numpy(), 0, i & 1)
        ts_submit.append(time.perf_counter() - t0)
    for i in range(N - 2, N):
        det.collect(i & 1)
        ts_collect.append(time.perf_counter() - t0)
d = [round((b - a) * 1e6) for a, b in zip(ts_collect, ts_collect[1:])]
print("collect-to-collect us:", d[:40])
print("mean us/step", round(sum(d) / len(d), 1), " submit call us:", [round((b - a) * 1e6) for a, b in zip(ts_collect[:10], ts_submit[2:12])])
