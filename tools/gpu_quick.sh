#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for v in "$@"; do
  eval "$v $B" > gpurun_out/q.json 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
print(sys.argv[1], "| sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "step_ms", round(d["ms_per_step"],3))
PY
done
