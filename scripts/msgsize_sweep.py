#!/usr/bin/env python
"""Device-resident detector throughput against the message size, for the kernel variants
(DM_KERNEL).  BASELINE config 2 fixes 64k x 256 B records per call (bench.py); this shows how
the decompositions behave when a caller batches more (or fewer) records per message."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import AuditSynth, MONITORED_KEYS


def main():
    variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["rows", "lanes"]
    sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16384, 65536, 262144, 1048576]
    varlen = len(sys.argv) > 3 and sys.argv[3] == "varlen"      # BASELINE config 5 record lengths (32..4096 B)
    g0 = AuditSynth(seed=5)
    train = (g0.batch_varlen(65536, inject=False) if varlen else g0.batch(65536, inject=False))[0]
    dev = torch.device("cuda:0")
    keys = [k.encode() for k in MONITORED_KEYS]
    out = {}
    g = AuditSynth(seed=6)
    for lines in sizes:
        n_msgs = max(2, min(8, (320 << 20) // (lines * 256)))       # > L2 in total where memory allows
        msgs = [(g.batch_varlen(lines, inject=True) if varlen else g.batch(lines, inject=True))[0] for _ in range(n_msgs)]
        bufs = []
        for m in [train] + msgs:
            t = torch.zeros(len(m) + 64, dtype=torch.uint8, device=dev)
            t[:len(m)] = torch.frombuffer(bytearray(m), dtype=torch.uint8).to(dev)
            bufs.append(t)
        cap = max(lines, 65536) + 16
        flags = torch.zeros(cap, dtype=torch.uint8, device=dev)
        scores = torch.zeros(cap, dtype=torch.float32, device=dev)
        for var in variants:
            os.environ["DM_KERNEL"] = var
            det = DeviceDetector(keys, max_batch_bytes=max(len(train), max(len(m) for m in msgs)) + 4096, max_lines=max(lines, 65536) + 16,
                                 table_log2_slots=16)
            st = torch.cuda.Stream(device=dev)
            steps = max(20, min(400, (1 << 26) // lines))
            with torch.cuda.stream(st):
                det.enqueue_device(bufs[0].data_ptr(), len(train), n_train_lines=65536, flags_ptr=flags.data_ptr(),
                                   scores_ptr=scores.data_ptr(), out_cap_lines=cap, stream=st.cuda_stream)
                for i in range(5):
                    j = 1 + i % n_msgs
                    det.enqueue_device(bufs[j].data_ptr(), len(msgs[j - 1]), 0, flags.data_ptr(), scores.data_ptr(), cap, st.cuda_stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                st.synchronize()
                e0.record(st)
                for i in range(steps):
                    j = 1 + i % n_msgs
                    det.enqueue_device(bufs[j].data_ptr(), len(msgs[j - 1]), 0, flags.data_ptr(), scores.data_ptr(), cap, st.cuda_stream)
                e1.record(st)
                st.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out.setdefault(var, {})[str(lines)] = {"ms_per_message": round(ms, 4), "G_lines_per_s": round(lines / ms / 1e6, 3),
                                                  "GBps_algorithmic": round((len(msgs[0]) + 5 * lines) / ms / 1e6, 1), "bytes_per_message": len(msgs[0]),
                                                  "anomalies_last": int(flags[:lines].sum().item())}
            det.close()
        del bufs
        torch.cuda.empty_cache()
    print(json.dumps({"workload": ("32..4096 B (config 5 mixture)" if varlen else "256 B") + " synthetic audit records, K=5 monitored fields, device-resident", "by_variant": out}))


if __name__ == "__main__":
    main()
