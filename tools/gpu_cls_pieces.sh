#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 240 > gpurun_out/cls_parity.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/cls_parity.log
run() {
  eval "$1 timeout 600 python bench.py $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e" > gpurun_out/q.json 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
  python - "$1 $2" <<'PY'
import json,sys
d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
print(sys.argv[1], "| sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "step_ms", round(d["ms_per_step"],3), "frac", round(r["frac"],4))
PY
}
run "MMQ_X=0" "--layout collapsed"
run "MMQ_DEBUG_CLS_SKIP=6" "--layout collapsed"
run "MMQ_DEBUG_CLS_SKIP=3" "--layout collapsed"
run "MMQ_CLS_GEO_HI=5" "--layout collapsed"
run "MMQ_CLS_GEO_LO=10" "--layout collapsed"
