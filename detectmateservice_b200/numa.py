"""NUMA placement of pinned host buffers.

A pinned buffer lives on the NUMA node of the thread that allocated it; when that is not the
node the GPU's PCIe root hangs off, host -> device copies run at about half the link rate
(measured on the B200 boxes: 16 MiB in ~560 us instead of ~311 us, `scripts/e2e_steps.py`).
`bound_to_gpu_node(device)` pins the calling thread to the GPU-local CPUs while the buffers
are allocated (and first touched) and restores the affinity afterwards.  Pure host-side
plumbing around the boundary; nothing here computes.
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional, Set


def parse_cpulist(text: str) -> Set[int]:
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11} (the format of /sys/devices/system/node/nodeN/cpulist)."""
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_cpus(device: int = 0, sysfs: str = "/sys") -> Optional[Set[int]]:
    """CPUs of the NUMA node the CUDA device `device` is attached to, None if unknown."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)) as f:
            cpus = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


@contextlib.contextmanager
def bound_to_gpu_node(device: int = 0):
    """Run the body with the calling thread bound to the GPU-local CPUs (no-op when the
    topology cannot be read).  Yields the CPU set used, or None."""
    cpus = gpu_numa_cpus(device)
    old = None
    if cpus:
        try:
            old = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
        except OSError:
            old, cpus = None, None
    try:
        yield cpus
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except OSError:
                pass
