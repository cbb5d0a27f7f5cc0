#!/bin/bash
# quick device-resident sweep of a tuning env var: scripts/sweep.sh VAR v1 v2 ...
VAR=$1; shift
for V in "$@"; do
  env $VAR=$V python - <<PY
import os,sys,time
sys.path.insert(0,'.')
import torch
from bench import _make_messages, LINES_PER_MSG
from detectmateservice_b200.detector import DeviceDetector
from detectmateservice_b200.synth import MONITORED_KEYS
msgs=_make_messages(0,n_msgs=8)
det=DeviceDetector(MONITORED_KEYS,max_batch_bytes=len(msgs[0])+4096,max_lines=LINES_PER_MSG+16,table_log2_slots=16)
d=[]
for m in msgs:
    t=torch.zeros(len(m)+64,dtype=torch.uint8,device='cuda'); t[:len(m)].copy_(torch.frombuffer(bytearray(m),dtype=torch.uint8)); d.append(t)
st=torch.cuda.Stream(); sp=st.cuda_stream
det.enqueue_device(d[0].data_ptr(),len(msgs[0]),LINES_PER_MSG,0,0,0,sp); det.sync()
for i in range(20): det.enqueue_device(d[1+i%7].data_ptr(),len(msgs[0]),0,0,0,0,sp)
det.sync(); det.profile_enable(True)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(100): det.enqueue_device(d[1+i%7].data_ptr(),len(msgs[0]),0,0,0,0,sp)
e1.record(st); torch.cuda.synchronize()
ms,n,_=det.profile_read()
print("$VAR=$V step %.2f us  main kernel %.2f us  anomalies %d"%(e0.elapsed_time(e1)*10, ms/n*1e3, det.sync()[1]))
PY
done
