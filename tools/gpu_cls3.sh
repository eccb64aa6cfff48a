#!/bin/bash
# class-plan kernel v3: geometry / prefetch sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 240 -k "class_plan or explicit_class or bit_exact_vs_cpu" > gpurun_out/cls_parity.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/cls_parity.log
B="python bench.py --layout collapsed --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for v in "MMQ_X=0" "MMQ_CLS_PREFETCH=0" "MMQ_CLS_GEO_LO=8" "MMQ_CLS_GEO_LO=12" "MMQ_CLS_GEO_HI=4" "MMQ_CLS_GEO_HI=6" "MMQ_DEBUG_CLS_SKIP=6" "MMQ_DEBUG_CLS_SKIP=3" "MMQ_DEBUG_CLS_SKIP=6 MMQ_CLS_PREFETCH=0" "MMQ_DEBUG_CLS_SKIP=3 MMQ_CLS_PREFETCH=0"; do
  eval "$v timeout 600 $B" > gpurun_out/q.json 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
print(sys.argv[1], "| sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "step_ms", round(d["ms_per_step"],3), "frac", round(r["frac"],4))
PY
done
