#!/usr/bin/env python
"""Where does the end-to-end (host buffer in, host results out) time go?  Run on a GPU box."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import _make_messages, LINES_PER_MSG
from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import MONITORED_KEYS

msgs = _make_messages(0, n_msgs=8)
det = DeviceDetector(MONITORED_KEYS, max_batch_bytes=len(msgs[0]) + 4096, max_lines=LINES_PER_MSG + 16, table_log2_slots=16)
pin = []
for m in msgs:
    t = torch.empty(len(m), dtype=torch.uint8, pin_memory=True)
    t.copy_(torch.frombuffer(bytearray(m), dtype=torch.uint8))
    pin.append(t)
det.process_lines(pin[0].numpy(), LINES_PER_MSG)
dev = torch.zeros(len(msgs[0]) + 64, dtype=torch.uint8, device="cuda")
N = 40


def timeit(name, fn):
    fn(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(N)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / N
    print(f"{name:48s} {dt*1e6:9.1f} us/step  {LINES_PER_MSG/dt/1e6:8.1f} M lines/s", flush=True)


def h2d_only(n):
    for i in range(n):
        dev[:len(msgs[0])].copy_(pin[1 + i % 7], non_blocking=True)


def sync_path(n):
    for i in range(n):
        det.process_lines(pin[1 + i % 7].numpy(), 0, copy=False)


def pipelined(n):
    for i in range(n):
        if i >= 2:
            det.collect(i & 1)
        det.submit(pin[1 + i % 7].numpy(), 0, i & 1)
    for i in range(max(0, n - 2), n):
        det.collect(i & 1)


def submit_only_then_collect(n):
    # how fast can submissions be issued if results are collected late (needs n <= 2 slots)?
    for i in range(0, n, 2):
        det.submit(pin[1].numpy(), 0, 0)
        det.submit(pin[2].numpy(), 0, 1)
        det.collect(0)
        det.collect(1)


def device_only(n):
    st = torch.cuda.Stream()
    for i in range(n):
        det.enqueue_device(dev.data_ptr(), len(msgs[0]), 0, 0, 0, 0, st.cuda_stream)
    det.sync()


dev[:len(msgs[1])].copy_(pin[1])
timeit("torch pinned H2D only", h2d_only)
timeit("device-resident kernels only", device_only)
timeit("dm_process_lines (sync, pinned in, pinned out)", sync_path)
timeit("dm_submit_lines/dm_collect, 2 slots", pipelined)
timeit("submit,submit,collect,collect", submit_only_then_collect)
# ctypes / python overhead of an empty message round trip
t0 = time.perf_counter()
for i in range(200):
    det.process_lines(b"", 0)
print(f"empty-message call overhead {(time.perf_counter()-t0)/200*1e6:.1f} us")
