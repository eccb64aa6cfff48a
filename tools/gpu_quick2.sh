#!/bin/bash
# like gpu_quick.sh but robust: one python run per variant, prints alloc_ms
mkdir -p gpurun_out
for v in "$@"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/q.json 2>gpurun_out/q.err
  python - "$v" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
    print(sys.argv[1], "| sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "step_ms", round(d["ms_per_step"],3))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/q.err").read()[-300:])
PY
done
