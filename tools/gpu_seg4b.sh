#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "by_length or categorical or multi_sweep or graph" > gpurun_out/pytest_seg4.log 2>&1; tail -3 gpurun_out/pytest_seg4.log
bash tools/gpu_quick2.sh "MMQ_SEG_GEO=0" "MMQ_SEG_ASCENDING=1" "MMQ_SEG_GEO=1" "MMQ_SEG_GEO=4"
for v in "MMQ_SEG_GEO=0" "MMQ_SEG_ASCENDING=1"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --weights > gpurun_out/q.json 2>gpurun_out/q.err
  python - "$v weighted" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
    print(sys.argv[1], "| sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "step_ms", round(d["ms_per_step"],3))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/q.err").read()[-300:])
PY
done
