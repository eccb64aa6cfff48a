#!/bin/bash
# privatised counts[] replicas: collapsed (class plan) and per-fragment (segment kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu --timeout 300 > gpurun_out/cls_parity.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/cls_parity.log
run() {
  eval "$1 timeout 600 python bench.py $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e" > gpurun_out/q.json 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
  python - "$1 $2" <<'PY'
import json,sys
d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
print(sys.argv[1], "| sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "step_ms", round(d["ms_per_step"],3), "frac", round(r["frac"],4))
PY
}
for R in 1 4 8 16; do run "MMQ_COUNT_REPLICAS=$R" "--layout collapsed"; done
for R in 1 8; do run "MMQ_COUNT_REPLICAS=$R MMQ_DEBUG_CLS_SKIP=6" "--layout collapsed"; run "MMQ_COUNT_REPLICAS=$R MMQ_DEBUG_CLS_SKIP=3" "--layout collapsed"; done
for R in 1 4 8 16; do run "MMQ_COUNT_REPLICAS=$R" ""; done
