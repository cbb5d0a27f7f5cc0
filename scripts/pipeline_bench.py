#!/usr/bin/env python
"""BASELINE config 3 measurement: the detector stage INSIDE the service plumbing.

    sender process --(NNG pair0, ipc|tcp)--> DetectorEngine(B200NewValueDetector, raw mode) --> sink

The sender streams M messages of `--lines` synthetic audit records; the engine thread does
recv -> process() -> send exactly like the reference loop (engine.py:163-246), so nothing
overlaps across messages.  Reports lines/s through the whole stage and where the time goes
(socket receive vs process()).  `--transport-only` swaps the detector for a processor that
returns None, which gives the transport ceiling of the pure-Python pynng shim on this host.

Not part of bench.py (whose e2e number is the C-ABI call with host buffers); this is the
"what does a user of the unmodified service see" figure quoted in DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_sender(addr: str, n_msgs: int, lines: int, pool: int) -> None:
    from detectmateservice_b200 import compat
    compat.install_shims()
    import pynng
    from detectmateservice_b200.synth import config2_stream
    msgs = [m for m, _ in config2_stream(n_lines=lines * pool, batch_lines=lines, train_lines=lines)]
    with pynng.Pair0(dial=addr, send_timeout=60000) as s:
        time.sleep(0.3)
        for i in range(n_msgs):
            s.send(msgs[0] if i == 0 else msgs[1 + (i - 1) % (pool - 1)])
        time.sleep(1.0)                                     # let the last frame drain before close


class _Null:
    """Transport-only processor: counts, detects nothing (NOT a detector fallback)."""

    def process(self, raw):
        return None


class _Timed:
    accepts_bytes_like = True

    def __init__(self, inner):
        self.inner, self.t_proc, self.n = inner, 0.0, 0
        self.n_seen = 0
        self.t_first = self.t_last = None

    def __getattr__(self, name):                            # alloc_frame / release_frame of the component
        if name in ("alloc_frame", "release_frame"):
            return getattr(self.inner, name)
        raise AttributeError(name)

    def process(self, raw):
        t0 = time.perf_counter()
        out = self.inner.process(raw)
        t1 = time.perf_counter()
        if self.t_first is None:
            self.t_first = t1                               # clock starts after message 0 (lazy CUDA init + training)
        self.n_seen = getattr(self.inner, "n_seen", self.n_seen + 1)
        if self.n > 0:                                      # message 0 = training + lazy init
            self.t_proc += t1 - t0
        self.n += 1
        self.t_last = t1
        return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--role", default="main")
    ap.add_argument("--addr", default="")
    ap.add_argument("--messages", type=int, default=64)
    ap.add_argument("--lines", type=int, default=65536)
    ap.add_argument("--pool", type=int, default=5)
    ap.add_argument("--transport", default="ipc", choices=["ipc", "tcp"])
    ap.add_argument("--output-format", default="compact", choices=["compact", "alerts"])
    ap.add_argument("--transport-only", action="store_true")
    ap.add_argument("--diag", action="store_true", help="time the shim's receive pieces (alloc vs socket reads)")
    a = ap.parse_args()
    if a.role == "sender":
        run_sender(a.addr, a.messages, a.lines, a.pool)
        return

    from detectmateservice_b200 import compat
    compat.install_shims()
    import pynng
    from detectmateservice_b200.service import DetectorEngine
    from detectmateservice_b200.synth import MONITORED_KEYS

    diag = {"alloc_s": 0.0, "recv_s": 0.0, "frames": 0, "recv_calls": 0}
    if a.diag:
        shim = sys.modules["pynng"]

        def timed_recv_into_new(sock, n):
            t0 = time.perf_counter()
            buf = bytearray(n)
            mv = memoryview(buf)
            t1 = time.perf_counter()
            got = 0
            while got < n:
                k = sock.recv_into(mv[got:], n - got)
                if k == 0:
                    return None
                got += k
                diag["recv_calls"] += 1
            diag["alloc_s"] += t1 - t0
            diag["recv_s"] += time.perf_counter() - t1
            diag["frames"] += 1
            return buf
        shim._recv_into_new = timed_recv_into_new

    tag = f"{os.getpid()}"
    if a.transport == "ipc":
        addr, out = f"ipc:///tmp/dm_pipe_{tag}.ipc", f"ipc:///tmp/dm_pipe_{tag}_out.ipc"
    else:
        addr, out = "tcp://127.0.0.1:47611", "tcp://127.0.0.1:47612"

    if a.transport_only:
        proc = _Timed(_Null())
    else:
        from detectmateservice_b200.component import B200NewValueDetector
        cfg = {"detectors": {"B200NewValueDetector": {
            "method_type": "new_value_detector", "data_use_training": a.lines, "auto_config": False,
            "global": {"g": {"header_variables": [{"pos": k} for k in MONITORED_KEYS]}},
            "params": {"input_format": "raw_lines", "output_format": a.output_format,
                       "max_batch_bytes": max(64 << 20, a.lines * 512)}}}}
        proc = _Timed(B200NewValueDetector(name="B200NewValueDetector", config=cfg))

    got = {"n": 0, "bytes": 0}
    sink = pynng.Pair0(listen=out, recv_timeout=200)
    stop = threading.Event()

    def drain():
        while not stop.is_set():
            try:
                m = sink.recv()
            except pynng.Timeout:
                continue
            except pynng.NNGException:
                break
            got["n"] += 1
            got["bytes"] += len(m)

    th = threading.Thread(target=drain, daemon=True)
    th.start()
    with DetectorEngine(proc, addr, out_addr=[out]) as eng:
        snd = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--role", "sender", "--addr", addr,
                                "--messages", str(a.messages), "--lines", str(a.lines), "--pool", str(a.pool)])
        t_end = time.monotonic() + 600
        while proc.n < a.messages and time.monotonic() < t_end and snd.poll() is None:
            time.sleep(0.01)
        t_end = time.monotonic() + 10
        while proc.n < a.messages and time.monotonic() < t_end:
            time.sleep(0.01)
        snd.wait(timeout=30)
        c = dict(eng.counters)
    stop.set()
    th.join(timeout=2)
    sink.close()
    n_timed = proc.n - 1
    wall = (proc.t_last - proc.t_first) if proc.n > 1 else float("nan")
    res = {
        "what": "transport-only" if a.transport_only else "detector stage in the service plumbing",
        "transport": a.transport, "messages": proc.n, "lines_per_message": a.lines,
        "bytes_per_message": c["read_bytes"] // max(1, c["messages"]),
        "lines_per_s": round(a.lines * n_timed / max(1e-9, wall), 1) if n_timed > 0 else None,
        "ms_per_message_wall": round(1e3 * wall / max(1, n_timed), 3),
        "ms_per_message_in_process": round(1e3 * proc.t_proc / max(1, n_timed), 3),
        "replies": got["n"], "reply_bytes": got["bytes"], "engine": c,
    }
    if a.diag and diag["frames"]:
        res["diag_ms_per_frame"] = {"alloc": round(1e3 * diag["alloc_s"] / diag["frames"], 3),
                                    "socket_reads": round(1e3 * diag["recv_s"] / diag["frames"], 3),
                                    "recv_calls": diag["recv_calls"] // diag["frames"]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
