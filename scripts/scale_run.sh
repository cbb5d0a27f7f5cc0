#!/bin/bash
# N = 1,2,4,8 back to back on one box (what the driver does at round end)
OUT=gpurun_out/${1:-scale}; mkdir -p $OUT
for N in ${NS:-1 2 4 8}; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 100 --warmup 10 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  fi
  echo "N=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n$N.json") if l.startswith("{")][-1])
    print("  value %.3e lines/s  ms/step %.4f  e2e %.3e  frac %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["roofline"]["frac"]))
except Exception as e: print("  parse error", e)
PY
done
