#!/usr/bin/env python
"""Which NUMA node do pinned buffers land on, and how fast is H2D from each?"""
import os, re, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectmateservice_b200.numa import gpu_numa_cpus
p = torch.cuda.get_device_properties(0)
bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
print("gpu", bdf, "numa_node", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip(), "local cpus", len(gpu_numa_cpus(0) or []))
for n in sorted(os.listdir("/sys/devices/system/node")):
    if n.startswith("node"):
        mi = open(f"/sys/devices/system/node/{n}/meminfo").read()
        tot = re.search(r"MemTotal:\s+(\d+)", mi).group(1); free = re.search(r"MemFree:\s+(\d+)", mi).group(1)
        print(n, "total GB", int(tot) >> 20, "free GB", int(free) >> 20, "cpus", open(f"/sys/devices/system/node/{n}/cpulist").read().strip())
try:
    print("mems_allowed", [l for l in open("/proc/self/status") if "Mems_allowed_list" in l or "Cpus_allowed_list" in l])
except Exception as e:
    print(e)
N = 16 << 20
import contextlib
from detectmateservice_b200.numa import bound_to_gpu_node
ctx = bound_to_gpu_node(0) if os.environ.get("BIND", "1") == "1" else contextlib.nullcontext()
print("BIND", os.environ.get("BIND", "1"))
with ctx:
    bufs = [torch.empty(N, dtype=torch.uint8, pin_memory=True) for _ in range(10)]
    for b in bufs:
        b.fill_(7)
if os.environ.get("FLUSH", "0") == "1":
    from detectmateservice_b200 import _lib
    L = _lib.load()
    for b in bufs:
        _lib.check(L.dm_host_cache_flush(b.data_ptr(), b.numel()))
    print("flushed")
maps = open("/proc/self/numa_maps").read().splitlines()
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
for i, b in enumerate(bufs):
    addr = b.data_ptr()
    line = ""
    for ln in maps:
        a = int(ln.split()[0], 16)
        if a <= addr:
            best = ln
        else:
            break
    nodes = re.findall(r"N\d+=\d+", best)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); dev.copy_(b, non_blocking=True); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e6)
    print(f"buf {i} addr {addr:#x} h2d us {min(ts):7.1f}  map {best.split()[0]} {' '.join(nodes)} {'huge' if 'huge' in best else ''}")
