#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench, ncu launch list + full capture of the
# dominant kernel.  Everything lands in gpurun_out/ (merged back by gpurun).
# usage: scripts/gpu_round.sh <tag> [stages...]   stages: smoke tests bench ncu sanitize
set -u
TAG=${1:-r01}
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
    tests_all)
      timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest_gpu.log" 2>&1
      echo "pytest_gpu rc=$?" | tee -a "$OUT/summary.txt"
      tail -15 "$OUT/pytest_gpu.log" ;;
    bench)
      for K in ${DM_KERNELS:-v1}; do
        DM_KERNEL=$K timeout 900 python bench.py > "$OUT/bench_$K.json" 2> "$OUT/bench_$K.err"
        echo "bench[$K] rc=$?" | tee -a "$OUT/summary.txt"
        cat "$OUT/bench_$K.json"
      done ;;
    ncu)
      for K in ${DM_KERNELS:-v1}; do
        DM_KERNEL=$K timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
          --log-file "$OUT/launches_$K.csv" python bench.py --steps 6 --warmup 3 --no-cpu > "$OUT/ncu_list_$K.log" 2>&1
        echo "ncu_list[$K] rc=$?" | tee -a "$OUT/summary.txt"
        DM_KERNEL=$K timeout 1200 ncu --set full --clock-control none ${NCU_EXTRA:-} --import-source on \
          -k regex:"${NCU_KERNEL:-dm_k_detect_lines|dm_k_tile|dm_k_rows}" -s 4 -c 3 -f -o "$OUT/prof_$K" \
          python bench.py --steps 6 --warmup 3 --no-cpu > "$OUT/ncu_full_$K.log" 2>&1
        echo "ncu_full[$K] rc=$?" | tee -a "$OUT/summary.txt"
      done ;;
    pipeline)
      for T in ipc tcp; do
        timeout 600 python scripts/pipeline_bench.py --transport $T --messages 48 > "$OUT/pipeline_$T.json" 2> "$OUT/pipeline_$T.err"
        echo "pipeline[$T] rc=$?" | tee -a "$OUT/summary.txt"
        cat "$OUT/pipeline_$T.json"
      done
      timeout 600 python scripts/pipeline_bench.py --transport ipc --messages 48 --output-format alerts > "$OUT/pipeline_ipc_alerts.json" 2> "$OUT/pipeline_ipc_alerts.err"
      echo "pipeline[ipc,alerts] rc=$?" | tee -a "$OUT/summary.txt"
      cat "$OUT/pipeline_ipc_alerts.json"
      timeout 600 python scripts/pipeline_bench.py --transport ipc --messages 48 --transport-only > "$OUT/pipeline_transport_only.json" 2>&1
      cat "$OUT/pipeline_transport_only.json" ;;
    racecheck)
      DM_KERNEL=${DM_KERNELS:-rows} timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 \
        python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/racecheck.log" 2>&1
      echo "racecheck rc=$?" | tee -a "$OUT/summary.txt"
      tail -5 "$OUT/racecheck.log" ;;
    sanitize)
      DM_KERNEL=${DM_KERNELS:-v1} timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 \
        python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/sanitize.log" 2>&1
      echo "sanitize rc=$?" | tee -a "$OUT/summary.txt"
      tail -5 "$OUT/sanitize.log" ;;
  esac
done
cat "$OUT/summary.txt"
