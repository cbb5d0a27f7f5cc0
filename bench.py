#!/usr/bin/env python
"""bench.py -- log lines/sec through the B200 detector (BASELINE.json metric).

A step = one pass of the hot path over one message of 65 536 synthetic audit records of
256 B (16 MiB).  Workload = BASELINE config 2 ("detector Service.process on 1xB200, 1M
synthetic 256B audit-log lines, batch 64k"): 16 distinct messages per GPU (256 MiB, larger
than the 126 MB L2, so successive steps never re-read a cached message); message 0 is the
anomaly-free training window (consumed untimed), the timed steps cycle over the 15
detection messages.  Multi-GPU: the stream shards by message across ranks (weak scaling,
every rank owns 16 messages), one NCCL all-reduce of the per-window statistics per step.

  value     lines/s, messages resident in HBM when the timed region starts, CUDA events.
  e2e       same metric through the C-ABI call with HOST buffers: pinned-host message in,
            H2D + kernels + D2H of flags/scores inside the timed region.
  roofline  dominant kernel (tokenizer+detector): algorithmic bytes (record bytes + 1 B flag
            + 4 B score per record) / its CUDA-event duration, against MEASURED_PEAKS.json.
  cpu_baseline  the oracle's C restatement (oracle/c/dm_oracle.c) on the host cores, bounded
            sample; N=1, rank 0 only.

`--impl reference` times that same CPU restatement with all host threads (the reference is
pure Python whose detector arithmetic is not vendored; see DESIGN.md "Reference arm").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LINES_PER_MSG = 65536
LINE_BYTES = 256
N_MSGS = 16
METRIC = "log lines/sec through detector"
UNIT = "lines/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _make_messages(seed_offset: int, n_msgs: int = N_MSGS, lines: int = LINES_PER_MSG):
    from detectmateservice_b200.synth import SEED, AuditSynth
    g = AuditSynth(SEED + seed_offset)
    msgs = [g.batch(lines, inject=False, line_bytes=LINE_BYTES)[0]]
    for _ in range(n_msgs - 1):
        msgs.append(g.batch(lines, inject=True, line_bytes=LINE_BYTES)[0])
    return msgs


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {}
        for attr, name in (("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap"),
                           ("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"),
                           ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"),
                           ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                           ("nvmlClocksThrottleReasonHwPowerBrakeSlowdown", "hw_power_brake")):
            if hasattr(nv, attr):
                names[getattr(nv, attr)] = name
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement, threads over shards of the same workload
# ------------------------------------------------------------------------------------------
def _cpu_oracles(msgs, threads: int):
    """One trained C-oracle instance per thread (training window = message 0, untimed)."""
    from detectmateservice_b200.synth import MONITORED_KEYS
    from oracle.native import NativeOracle
    keys = [k.encode() for k in MONITORED_KEYS]
    oracles = []
    for _ in range(threads):
        o = NativeOracle(keys)
        o.process(msgs[0], LINES_PER_MSG)
        oracles.append(o)
    return oracles


def _cpu_throughput(msgs, sample_lines: int, threads: int, repeats: int = 1, oracles=None):
    """Detect `sample_lines` records per thread with `threads` threads (each owns a trained
    oracle instance).  Returns (lines/s, seconds)."""
    if oracles is None:
        oracles = _cpu_oracles(msgs, threads)
    nbytes = sample_lines * LINE_BYTES
    shards = [np.frombuffer(msgs[1 + (t % (len(msgs) - 1))], dtype=np.uint8)[:nbytes] for t in range(threads)]
    barrier = threading.Barrier(threads + 1)

    def work(t):
        barrier.wait()
        for _ in range(repeats):
            oracles[t].process(shards[t], 0)
        barrier.wait()

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for th in ths:
        th.join()
    return threads * sample_lines * repeats / dt, dt


def _python_per_record_rate(msgs, n_lines: int = 4000):
    """The reference's granularity: one record per process() call in pure Python (oracle/nvd.py)."""
    from detectmateservice_b200.synth import MONITORED_KEYS
    from oracle.nvd import NewValueDetectorOracle
    from oracle import rtok
    cfg = {"data_use_training": 2000, "global": {"g": {"header_variables": [{"pos": k} for k in MONITORED_KEYS]}}}
    det = NewValueDetectorOracle(config=cfg)
    lines = rtok.split_records(msgs[0][:LINE_BYTES * 2000]) + rtok.split_records(msgs[1][:LINE_BYTES * n_lines])
    for l in lines[:2000]:
        det.step_line(l)
    t0 = time.perf_counter()
    for l in lines[2000:]:
        det.step_line(l)
    return n_lines / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import native
    native.build()
    threads = os.cpu_count() or 1
    msgs = _make_messages(0, n_msgs=4)
    # bounded sample per step: 8192 records per thread (about 2 MiB each)
    sample = 8192
    oracles = _cpu_oracles(msgs, threads)
    for _ in range(max(min(args.warmup, 5), 1)):
        _cpu_throughput(msgs, sample, threads, oracles=oracles)
    t_total, lines_total = 0.0, 0
    for _ in range(args.steps):
        rate, dt = _cpu_throughput(msgs, sample, threads, oracles=oracles)
        t_total += dt
        lines_total += sample * threads
    value = lines_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "config2: 64k x 256 B synthetic audit records per message, K=5 monitored fields",
                   "lines_per_message": LINES_PER_MSG, "line_bytes": LINE_BYTES},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} records per thread per step, {threads} threads, C restatement "
                                   f"oracle/c/dm_oracle.c (reference detector source is not vendored)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from detectmateservice_b200 import _lib
    from detectmateservice_b200.detector import DeviceDetector
    from detectmateservice_b200.synth import MONITORED_KEYS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line (NCCL_DEBUG=VERSION/INFO prints a banner there)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("DM_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    msgs = _make_messages(1000 * rank)
    nbytes = [len(m) for m in msgs]
    n_lines_msg = [m.count(b"\n") for m in msgs]
    det = DeviceDetector(MONITORED_KEYS, device=local_rank, max_batch_bytes=max(nbytes) + 4096,
                         max_lines=LINES_PER_MSG + 16, table_log2_slots=16)
    # a dedicated non-default stream: the C ABI treats a NULL stream as "the handle's own
    # stream", and torch's default stream IS NULL -- events must sit on the launching stream.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    # device-resident copies (value) and pinned-host copies (e2e)
    # (pinned buffers are allocated with the thread bound to the GPU's NUMA node: a buffer on the
    # other socket halves the host->device rate, detectmateservice_b200/numa.py)
    from detectmateservice_b200.numa import bound_to_gpu_node
    d_msgs, h_msgs = [], []
    with bound_to_gpu_node(local_rank) as numa_cpus:
        for m in msgs:
            t = torch.zeros(len(m) + 64, dtype=torch.uint8, device=dev)
            src = torch.frombuffer(bytearray(m), dtype=torch.uint8)
            t[:len(m)].copy_(src)
            d_msgs.append(t)
            hp = torch.empty(len(m), dtype=torch.uint8, pin_memory=True)
            hp.copy_(src)
            h_msgs.append(hp)
    # the set-up just wrote these buffers: evict them from the CPU caches, else the H2D copies
    # of whichever buffers are still cached run at a fraction of the link rate (dmdetect.cu)
    from detectmateservice_b200 import _lib as _dmlib
    for hp in h_msgs:
        _dmlib.check(_dmlib.load().dm_host_cache_flush(hp.data_ptr(), hp.numel()))
    d_flags = torch.zeros(LINES_PER_MSG + 16, dtype=torch.uint8, device=dev)
    d_scores = torch.zeros(LINES_PER_MSG + 16, dtype=torch.float32, device=dev)
    cap = LINES_PER_MSG + 16
    from detectmateservice_b200.window import DeviceWindow
    dwin = DeviceWindow(det, rank, world, dev)
    if world > 1 and os.environ.get("DM_WINDOW", "native") == "native":
        # the library's own NCCL communicator: one C call per window (export, ncclAllReduce, import);
        # DM_WINDOW=torch keeps the torch.distributed.all_reduce route
        dwin.init_native(n_comms=2 if os.environ.get("DM_WINDOW_COMMS", "2") == "2" else 1)

    # the per-window exchange runs on a side stream: in steady state it carries statistics
    # only and gates nothing, so it overlaps the next message's kernels
    side = torch.cuda.Stream(device=dev)
    sides = [side, torch.cuda.Stream(device=dev)]
    win_n = [0]
    n_sides = 2 if os.environ.get("DM_WINDOW_COMMS", "2") == "2" else 1
    win_ev = torch.cuda.Event()

    def window(with_keys: bool):
        if with_keys:
            dwin.exchange(True, sp)                 # training window: detection must wait for it
            return
        sd = sides[win_n[0] % n_sides]               # DM_WINDOW_COMMS=2: two windows' all-reduces in flight
        win_n[0] += 1
        win_ev.record(stream)
        sd.wait_event(win_ev)
        if getattr(dwin, "native", False):
            dwin.exchange(False, sd.cuda_stream)
        else:
            with torch.cuda.stream(sd):              # the torch.distributed route reduces on the current stream
                dwin.exchange(False, sd.cuda_stream)

    # training window (untimed): every rank learns its message 0, then one exchange with keys
    det.enqueue_device(d_msgs[0].data_ptr(), nbytes[0], n_lines_msg[0], d_flags.data_ptr(), d_scores.data_ptr(), cap, sp)
    window(True)
    det.sync()

    def step(i: int):
        j = 1 + (i % (N_MSGS - 1))
        det.enqueue_device(d_msgs[j].data_ptr(), nbytes[j], 0, d_flags.data_ptr(), d_scores.data_ptr(), cap, sp)
        if world > 1:
            window(False)
        return j

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events on the launching stream ----------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    _, _, launches0 = det.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lines_timed = 0
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        lines_timed += n_lines_msg[step(i)]
    stream.wait_stream(sides[0])                     # the last windows' all-reduces are part of the job
    stream.wait_stream(sides[1])
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.finish()
    _, _, launches1 = det.profile_read()
    n_anom_last = det.sync()[1]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(lines_timed)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, lines_all = float(t.item()), float(tot.item())
    value = lines_all / (ms_max * 1e-3)

    # ---- roofline: dominant kernel, event pair inside the library --------------------------
    det.profile_enable(True)
    alg_bytes = 0
    n_prof = min(args.steps, 60)
    for i in range(n_prof):
        j = step(i)
        alg_bytes += nbytes[j] + 5 * n_lines_msg[j]
    k_ms, k_n, _ = det.profile_read()
    det.profile_enable(False)
    peak, peak_src = _peaks()
    achieved = (alg_bytes / max(k_n, 1)) / (k_ms / max(k_n, 1) * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": k_ms / max(k_n, 1),
                "algorithmic_bytes_per_launch": alg_bytes / max(k_n, 1),
                "kernel": os.environ.get("DM_KERNEL", "default"),
                # the whole step (K_A + K_B, overlapped by programmatic dependent launch) on this rank
                "step_achieved": (alg_bytes / max(n_prof, 1)) / (ms / max(args.steps, 1) * 1e-3) / 1e9,
                "step_frac": (alg_bytes / max(n_prof, 1)) / (ms / max(args.steps, 1) * 1e-3) / 1e9 / peak}

    # ---- e2e: C-ABI calls with pinned HOST buffers, H2D + D2H inside the timed region --------
    # dm_submit_lines / dm_collect (two slots): message i+1 crosses PCIe while message i runs;
    # every step's flags and scores are read back into host memory.
    torch.cuda.synchronize()
    e2e_steps = max(4, min(args.steps, 240))
    pipelined = os.environ.get("DM_KERNEL", "rows") == "rows"

    def e2e_loop(n_steps: int) -> int:
        lines = 0
        if pipelined:
            for i in range(n_steps):
                slot = i & 1
                if i >= 2:
                    f, s = det.collect(slot)
                    lines += f.size
                det.submit(h_msgs[1 + (i % (N_MSGS - 1))].numpy(), 0, slot)
            for i in range(max(0, n_steps - 2), n_steps):
                f, s = det.collect(i & 1)
                lines += f.size
        else:
            for i in range(n_steps):
                f, s = det.process_lines(h_msgs[1 + (i % (N_MSGS - 1))].numpy(), 0, copy=False)
                lines += f.size
        return lines

    with bound_to_gpu_node(local_rank):              # the library's pinned result buffers are created here
        e2e_loop(4)
    barrier()
    t0 = time.perf_counter()
    e2e_lines = e2e_loop(e2e_steps)
    if world > 1:
        window(False)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(e2e_lines)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    e2e_value = float(tot.item()) / float(t.item())
    # plain H2D bandwidth of the same pinned buffers, for context
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(8):
        d_msgs[1 + i][:nbytes[1 + i]].copy_(h_msgs[1 + i], non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = sum(nbytes[1:9]) / (time.perf_counter() - t0) / 1e9
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes[1],
           "d2h_bytes_per_step": 5 * n_lines_msg[1] + 32, "steps": e2e_steps,
           "api": ("dm_submit_lines/dm_collect (2 slots, pinned host buffers)" if pipelined else
                   "dm_process_lines(host pinned buffer)") + " via DeviceDetector",
           "ms_per_step": 1e3 * float(t.item()) / e2e_steps, "h2d_gbs_pinned": h2d_gbs,
           "pinned_numa_local_cpus": len(numa_cpus) if numa_cpus else None}

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import native
        native.build()
        cores = os.cpu_count() or 1
        _cpu_throughput(msgs, 4096, cores)
        rate, secs = _cpu_throughput(msgs, 32768, cores, repeats=4)
        one, _ = _cpu_throughput(msgs, 32768, 1, repeats=2)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"4 x 32768 records per thread on {cores} threads ({secs:.2f} s), oracle/c/dm_oracle.c; "
                         f"reference detector (detectmatelibrary) is not vendored",
               "single_thread": one,
               "python_per_record": _python_per_record_rate(msgs)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config2: detector on 64k x 256 B synthetic audit records per message "
                                   "(1M records = 16 messages per GPU), K=5 monitored fields, p=1e-3 anomalies",
                       "lines_per_message": LINES_PER_MSG, "line_bytes": LINE_BYTES, "messages_per_gpu": N_MSGS,
                       "l2_policy": "inputs larger than L2 (15 x 16 MiB cycled per GPU)",
                       "parallelism": f"shard-by-message x{world}, 1 stats all-reduce per step" if world > 1 else "single GPU"},
            "e2e": e2e, "gpu_launches": int(launches1 - launches0), "clocks": clocks, "roofline": roofline,
            "anomalies_last_message": int(n_anom_last),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    det.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
