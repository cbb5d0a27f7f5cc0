#!/usr/bin/env python
"""Device-resident timing of the raw key=value path for a list of kernel configurations.
Development tool (gpurun): prints one JSON line per configuration.
usage: python scripts/stream_bench.py [--steps N] [--lines L] cfg ...   cfg = KERNEL[:overlap][:ctas]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--lines", type=int, default=65536)
    ap.add_argument("--msgs", type=int, default=16)
    ap.add_argument("--varlen", action="store_true")
    ap.add_argument("cfgs", nargs="*", default=["stream:1", "stream:0", "lanes:0"])
    args = ap.parse_args()
    import torch
    from detectmateservice_b200.detector import DeviceDetector
    from detectmateservice_b200.synth import MONITORED_KEYS, AuditSynth, SEED
    g = AuditSynth(SEED + (3 if args.varlen else 0))
    gen = g.batch_varlen if args.varlen else g.batch
    msgs = [gen(args.lines, inject=False)[0]] + [gen(args.lines, inject=True)[0] for _ in range(args.msgs - 1)]
    dev = torch.device("cuda", 0)
    d_msgs = []
    for m in msgs:
        t = torch.zeros(len(m) + 64, dtype=torch.uint8, device=dev)
        t[:len(m)].copy_(torch.frombuffer(bytearray(m), dtype=torch.uint8))
        d_msgs.append(t)
    cap = args.lines + 16
    d_flags = torch.zeros(cap, dtype=torch.uint8, device=dev)
    d_scores = torch.zeros(cap, dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    for cfg in args.cfgs:
        parts = cfg.split(":")
        os.environ["DM_KERNEL"] = parts[0]
        ov = int(parts[1]) if len(parts) > 1 else 0
        if len(parts) > 2:
            os.environ["DM_STREAM_CTAS_PER_SM"] = parts[2]
        else:
            os.environ.pop("DM_STREAM_CTAS_PER_SM", None)
        det = DeviceDetector(MONITORED_KEYS, device=0, max_batch_bytes=max(len(m) for m in msgs) + 4096,
                             max_lines=cap, table_log2_slots=16)
        det.set_overlap(bool(ov))
        det.enqueue_device(d_msgs[0].data_ptr(), len(msgs[0]), args.lines, d_flags.data_ptr(), d_scores.data_ptr(), cap, sp)
        det.sync()
        def step(i):
            j = 1 + (i % (args.msgs - 1))
            det.enqueue_device(d_msgs[j].data_ptr(), len(msgs[j]), 0, d_flags.data_ptr(), d_scores.data_ptr(), cap, sp)
            return j
        for i in range(20):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nb = 0
        for i in range(args.steps):
            nb += len(msgs[step(i)])
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        n_lines, n_anom = det.sync()
        st = det.stats()
        us = 1e3 * ms / args.steps
        print(json.dumps({"cfg": cfg, "us_per_step": round(us, 2), "lines_per_s": args.lines * args.steps / (ms * 1e-3),
                          "GBps": nb / (ms * 1e-3) / 1e9, "last_anomalies": n_anom, "anomalies_total": st["anomalies"],
                          "lines_total": st["lines"]}), flush=True)
        det.close()


if __name__ == "__main__":
    main()
