#!/bin/bash
# One GPU-box visit: smoke, GPU tests, bench (both arms), ncu launch list + one full capture of the
# dominant kernel, compute-sanitizer.  Everything lands in gpurun_out/<tag>/ (merged back by gpurun).
# usage: scripts/gpu_round.sh <tag> [stages...]   stages: smoke tests bench refarm ncu sanitize extras
set -u
TAG=${1:-r02}
shift || true
STAGES=${*:-smoke tests bench ncu}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.csv" 2>&1

for S in $STAGES; do
  case $S in
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
      echo "smoke rc=$?" | tee -a "$OUT/summary.txt" ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > "$OUT/pytest_gpu.log" 2>&1
      echo "pytest_gpu rc=$?" | tee -a "$OUT/summary.txt"
      tail -5 "$OUT/pytest_gpu.log" ;;
    bench)
      timeout 900 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
      echo "bench rc=$?" | tee -a "$OUT/summary.txt"
      cat "$OUT/bench_n1.json" ;;
    refarm)
      timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > "$OUT/bench_reference_n1.json" 2> "$OUT/bench_reference_n1.err"
      echo "refarm rc=$?" | tee -a "$OUT/summary.txt"
      cat "$OUT/bench_reference_n1.json" ;;
    ncu)
      # launch list of the bench command (cold-cache, serialised: shares, not absolutes) ...
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file "$OUT/launches.csv" python bench.py --steps 6 --warmup 3 --no-cpu --no-extra > "$OUT/ncu_list.log" 2>&1
      echo "ncu_list rc=$?" | tee -a "$OUT/summary.txt"
      # ... and one full capture of the dominant kernel (device-resident loop, 3 launches after warm-up)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dm_k_stream -s 30 -c 3 -f -o "$OUT/stream" \
        python scripts/stream_bench.py --steps 40 stream:0 > "$OUT/ncu_full.log" 2>&1
      echo "ncu_full rc=$?" | tee -a "$OUT/summary.txt" ;;
    sanitize)
      for M in thread chain; do
        DM_STREAM_RECHECK=$M timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 \
          python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/memcheck_$M.log" 2>&1
        echo "memcheck[$M] rc=$?" | tee -a "$OUT/summary.txt"
        tail -3 "$OUT/memcheck_$M.log"
      done ;;
    extras)
      timeout 300 python scripts/records_bench.py > "$OUT/records_bench.json" 2> "$OUT/records_bench.err"
      timeout 300 python scripts/format_bench.py > "$OUT/format_bench.json" 2> "$OUT/format_bench.err"
      timeout 200 python scripts/stream_bench.py stream:1 stream:0 lanes:0 > "$OUT/stream_bench_config2.jsonl" 2>&1
      timeout 300 python scripts/stream_bench.py --varlen stream:1 stream:0 lanes:0 > "$OUT/stream_bench_config5.jsonl" 2>&1
      DM_STREAM_RECHECK=thread timeout 300 python scripts/stream_bench.py --varlen stream:1 > "$OUT/stream_bench_config5_onebyone.jsonl" 2>&1
      timeout 120 python scripts/sock_ceiling.py > "$OUT/sock_ceiling.json" 2>&1
      echo "extras done" | tee -a "$OUT/summary.txt" ;;
  esac
done
cat "$OUT/summary.txt"
