#!/usr/bin/env python
"""What a unix-domain stream socket moves between two processes on this host, with nothing but a recv_into() loop
into one preallocated buffer on the receiving side -- the ceiling of ANY SP/PAIR0-over-ipc receive path, in C or in
Python (BASELINE config 3).  Prints one JSON line."""
import json
import multiprocessing as mp
import os
import socket
import time

N, M = 16 << 20, 48


def sender(path, sndbuf):
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    if sndbuf:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, sndbuf)
    s.connect(path)
    data = memoryview(bytearray(os.urandom(1 << 20)) * 16)
    for _ in range(M):
        s.sendall(data)
    s.close()


def run(bufsize):
    path = f"/tmp/dm_sock_ceiling_{os.getpid()}.sock"
    try:
        os.unlink(path)
    except OSError:
        pass
    ls = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    ls.bind(path)
    ls.listen(1)
    p = mp.Process(target=sender, args=(path, bufsize))
    p.start()
    c, _ = ls.accept()
    if bufsize:
        c.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, bufsize)
    buf = memoryview(bytearray(N))
    calls = 0
    t0 = time.perf_counter()
    for _ in range(M):
        got = 0
        while got < N:
            got += c.recv_into(buf[got:], N - got)
            calls += 1
    dt = time.perf_counter() - t0
    p.join()
    os.unlink(path)
    return {"GBps": round(N * M / dt / 1e9, 2), "recv_calls_per_16MiB": round(calls / M, 1),
            "records_per_s_at_256B": round(N * M / dt / 256, 0)}


if __name__ == "__main__":
    print(json.dumps({"what": "unix stream socket, sendall -> recv_into, 16 MiB messages, two processes",
                      "default_buffers": run(0), "buffers_1MiB": run(1 << 20), "buffers_4MiB": run(4 << 20)}))
