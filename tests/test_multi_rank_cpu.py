"""The N>1 path's host logic on CPU: world_size-2 gloo, record-aligned sharding, the
sum-all-reduce window buffer (statistics + learnt keys), against a single-rank oracle run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from detectmateservice_b200 import window
from oracle import fingerprint
from oracle.native import NativeOracle

KEYS = [b"type", b"exe", b"terminal", b"acct", b"res"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_are_record_aligned():
    from detectmateservice_b200.synth import AuditSynth
    msg, _ = AuditSynth(seed=1).batch_varlen(3000, inject=False)
    for world in (1, 2, 3, 8):
        cuts = window.shard_bounds(msg, world)
        assert cuts[0] == 0 and cuts[-1] == len(msg) and cuts == sorted(cuts) and len(cuts) == world + 1
        for c in cuts[1:-1]:
            assert c == len(msg) or msg[c - 1:c] == b"\n"
        assert b"".join(msg[cuts[i]:cuts[i + 1]] for i in range(world)) == msg
    assert window.shard_bounds(b"", 4) == [0, 0, 0, 0, 0]
    assert window.shard_bounds(b"no newline at all", 2) == [0, 17, 17]


def _rank_main(rank, world, port, train, detect, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank learns from ITS shard of the training window, then one exchange with keys
        cuts = window.shard_bounds(train, world)
        mine = NativeOracle(KEYS)
        shard = train[cuts[rank]:cuts[rank + 1]]
        n_train = shard.count(b"\n")
        mine.process(shard, n_train)
        keys = sorted(fingerprint.table_key(i, v) for i in range(len(KEYS)) for v in mine.known_values(i))
        stats = [n_train, n_train, 0, 0, 0, len(shard)]
        buf = torch.from_numpy(window.pack_window_host(stats, keys, rank, world, True))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        g_stats, per_rank = window.unpack_window_host(buf.numpy(), world, True)
        merged = sorted(set(int(k) for ks in per_rank for k in ks))
        # detection window: shard by record, statistics-only exchange
        dcuts = window.shard_bounds(detect, world)
        dshard = detect[dcuts[rank]:dcuts[rank + 1]]
        # (a rank detects against the MERGED set: emulate by training a fresh oracle on everything)
        full = NativeOracle(KEYS)
        full.process(train, train.count(b"\n"))
        f, s, _ = full.process(dshard, 0)
        dstats = [f.size, 0, f.size, int(f.sum()), int(s.sum()), len(dshard)]
        sbuf = torch.from_numpy(window.pack_window_host(dstats, [], rank, world, False))
        dist.all_reduce(sbuf, op=dist.ReduceOp.SUM)
        d_stats, _ = window.unpack_window_host(sbuf.numpy(), world, False)
        out_q.put((rank, g_stats[:6].tolist(), merged, [len(k) for k in per_rank], f.tolist(), d_stats[:6].tolist()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_window_exchange():
    from detectmateservice_b200.synth import AuditSynth
    g = AuditSynth(seed=42)
    train, _ = g.batch(4000, inject=False)
    detect, _ = g.batch(6000, inject=True)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, train, detect, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank ground truth
    one = NativeOracle(KEYS)
    one.process(train, 4000)
    want_keys = sorted(fingerprint.table_key(i, v) for i in range(len(KEYS)) for v in one.known_values(i))
    wf, ws, _ = one.process(detect, 0)
    assert res[0][1] == res[1][1] == [4000, 4000, 0, 0, 0, len(train)]          # summed training statistics
    assert res[0][2] == res[1][2] == want_keys                                   # union of the ranks' keys
    assert sum(res[0][3]) >= len(want_keys)                                      # segments concatenated, not added
    assert res[0][4] + res[1][4] == wf.tolist()                                  # rank-order concatenation
    assert res[0][5] == res[1][5] == [6000, 0, 6000, int(wf.sum()), int(ws.sum()), len(detect)]


def test_window_buffer_layout_limits():
    with pytest.raises(ValueError):
        window.pack_window_host([0] * 6, list(range(1, window.WINDOW_KEYS + 2)), 0, 2, True)
    assert window.window_words(8, False) == 40 and window.window_words(8, True) == 40 + 8 * (1 + window.WINDOW_KEYS)
    big = (1 << 63) + 12345                                                      # keys use all 64 bits
    buf = window.pack_window_host([1], [big], 1, 2, True) + window.pack_window_host([2], [7], 0, 2, True)
    st, per = window.unpack_window_host(buf, 2, True)
    assert st[0] == 3 and per[0].tolist() == [7] and per[1].tolist() == [big]
