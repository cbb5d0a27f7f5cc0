#!/usr/bin/env python
"""Where does the time of B200NewValueDetector.process() go for a 16 MiB pinned message?  (development tool)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import _make_messages, _component, LINES_PER_MSG
msgs = _make_messages(0, n_detect=4)
h = []
for m in msgs:
    t = torch.empty(len(m), dtype=torch.uint8, pin_memory=True); t.copy_(torch.frombuffer(bytearray(m), dtype=torch.uint8)); h.append(t)
for piece in (0, 2, 4, 8):
    comp = _component(0, len(msgs[0]) + 4096)
    if piece == 0:
        comp.PIPE_MIN_BYTES = 1 << 40
    else:
        comp.PIPE_PIECE_BYTES = piece << 20
    comp.process(memoryview(h[0].numpy()))
    for i in range(4):
        comp.process(memoryview(h[1 + i % 4].numpy()))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 24
    for i in range(n):
        comp.process(memoryview(h[1 + i % 4].numpy()))
    dt = (time.perf_counter() - t0) / n
    print(f"piece {piece} MiB: {dt*1e3:.3f} ms per 16 MiB message, {LINES_PER_MSG/dt/1e6:.1f} M lines/s", flush=True)
    # inside: submit / collect costs
    if piece:
        det = comp.det
        arr = h[1].numpy()
        t0 = time.perf_counter(); det.submit(arr[:piece << 20], 0, 0); t1 = time.perf_counter(); det.collect(0); t2 = time.perf_counter()
        print(f"   one piece alone: submit {1e6*(t1-t0):.0f} us, collect {1e6*(t2-t1):.0f} us")
    comp.close()
