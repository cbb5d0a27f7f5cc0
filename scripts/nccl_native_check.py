#!/usr/bin/env python
"""torchrun --nproc-per-node 2 scripts/nccl_native_check.py
The library's own window all-reduce (dm_window_allreduce) against the torch.distributed route:
two ranks train on different shards, exchange keys, then detect; global statistics and the
detection results of both routes must be identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from detectmateservice_b200 import window
from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import AuditSynth, MONITORED_KEYS

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl")
g = AuditSynth(seed=21)
train, _ = g.batch(6000, inject=False)
detect, _ = g.batch(6000, inject=True)
tc, dc = window.shard_bounds(train, world), window.shard_bounds(detect, world)
my_train, my_detect = train[tc[rank]:tc[rank + 1]], detect[dc[rank]:dc[rank + 1]]
keys = [k.encode() for k in MONITORED_KEYS]
res = {}
for route in ("torch", "native"):
    det = DeviceDetector(keys, device=dev.index, max_batch_bytes=4 << 20, table_log2_slots=14)
    w = window.DeviceWindow(det, rank, world, dev)
    if route == "native":
        w.init_native()
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        det.process_lines(my_train, n_train_lines=my_train.count(b"\n"))
        w.exchange(True, st.cuda_stream)
        f, s = det.process_lines(my_detect, 0)
        w.exchange(False, st.cuda_stream)
        st.synchronize()
    res[route] = (f.copy(), s.copy(), det.global_stats(), det.stats()["known_keys"])
    det.close()
ok = (res["torch"][0] == res["native"][0]).all() and (res["torch"][1] == res["native"][1]).all()
ok = ok and res["torch"][2] == res["native"][2] and res["torch"][3] == res["native"][3]
gs = res["native"][2]
print(f"rank {rank}: identical={bool(ok)} global lines={gs['lines']} anomalies={gs['anomalies']} known_keys={res['native'][3]}", flush=True)
t = torch.tensor([1 if ok and gs["lines"] == 12000 else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
