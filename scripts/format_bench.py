#!/usr/bin/env python
"""Throughput of the log_format + template mode (dm_kernels_format.cuh) on BASELINE config-2
shaped input (64k x 256 B audit records per message), device-resident, CUDA events on the
launching stream.  Not the headline metric (bench.py measures the key=value path); quoted in
DESIGN.md next to it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from detectmateservice_b200.component import parse_monitors, select_component_config
from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import AuditSynth

AUDIT = "type=<type> msg=audit(<Time>): <Content>"
TMPL = ("pid=<*> uid=<*> auid=<*> ses=<*> msg='op=<*> acct=<*> exe=<*> hostname=<*> addr=<*> terminal=<*> res=<*>' pad=<*>")
DECOYS = ["pid=<*> uid=<*> auid=<*> ses=<*> msg='unit=<*> comm=<*> exe=<*> hostname=<*> addr=<*> terminal=<*> res=<*>",
          "pid=<*> uid=<*> old-auid=<*> auid=<*> tty=<*> old-ses=<*> ses=<*> res=<*>",
          "apparmor=<*> operation=<*> info=<*> profile=<*> name=<*> pid=<*> comm=<*>"]


def main():
    steps, warmup, n_msgs, lines = 200, 20, 8, 65536
    cfg = {"detectors": {"NewValueDetector": {"method_type": "new_value_detector", "data_use_training": lines,
           "global": {"g": {"header_variables": [{"pos": "type"}]}},
           "events": {3: {"pam": {"variables": [{"pos": 5, "name": "acct"}, {"pos": 6, "name": "exe"},
                                                {"pos": 9, "name": "terminal"}, {"pos": 10, "name": "res"}]}}}}}}
    mons = parse_monitors(select_component_config(cfg, "NewValueDetector"))
    g = AuditSynth(seed=7)
    msgs = [g.batch(lines, inject=(i > 0))[0] for i in range(n_msgs)]
    dev = torch.device("cuda:0")
    bufs = []
    for m in msgs:
        t = torch.zeros(len(m) + 64, dtype=torch.uint8, device=dev)
        t[:len(m)] = torch.frombuffer(bytearray(m), dtype=torch.uint8).to(dev)
        bufs.append(t)
    flags = torch.zeros(lines, dtype=torch.uint8, device=dev)
    scores = torch.zeros(lines, dtype=torch.float32, device=dev)
    out = {}
    for label, templates in (("header+1 template", [TMPL]), ("header+4 templates (3 decoys first)", DECOYS + [TMPL])):
        det = DeviceDetector([m.key for m in mons], max_batch_bytes=32 << 20)
        mm = [{"event_id": (len(templates) - 1 if m.event_id is not None else None), "source": m.source, "pos": m.pos} for m in mons]
        det.set_monitors(mm)
        det.set_format(AUDIT, templates)
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            det.enqueue_device(bufs[0].data_ptr(), len(msgs[0]), n_train_lines=lines, flags_ptr=flags.data_ptr(),
                               scores_ptr=scores.data_ptr(), out_cap_lines=lines, stream=st.cuda_stream)
            for i in range(warmup):
                j = 1 + i % (n_msgs - 1)
                det.enqueue_device(bufs[j].data_ptr(), len(msgs[j]), 0, flags.data_ptr(), scores.data_ptr(), lines, st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st.synchronize()
            e0.record(st)
            for i in range(steps):
                j = 1 + i % (n_msgs - 1)
                det.enqueue_device(bufs[j].data_ptr(), len(msgs[j]), 0, flags.data_ptr(), scores.data_ptr(), lines, st.cuda_stream)
            e1.record(st)
            st.synchronize()
        ms = e0.elapsed_time(e1) / steps
        n_anom = int(flags.sum().item())
        out[label] = {"ms_per_message": round(ms, 4), "lines_per_s": round(lines / ms * 1e3, 1),
                      "GBps_algorithmic": round(lines * 261 / ms / 1e6, 2), "anomalies_last_message": n_anom}
        det.close()
    print(json.dumps({"workload": "64k x 256 B audit records per message, log_format + templates on the device", **out}))


if __name__ == "__main__":
    main()
