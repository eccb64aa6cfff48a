#!/bin/bash
mkdir -p gpurun_out
MMQ_CREATE_TIMING=1 python tools/time_create2.py 2>&1 | grep -v "iteration 0" | tail -44
MMQ_CREATE_TIMING=1 python bench.py --no-cpu-baseline > gpurun_out/q.json 2>gpurun_out/q.err; grep mmq_create gpurun_out/q.err | tail -16
python - <<'PY'
import json
d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
print("perfragment | sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "e2e", d["e2e"])
PY
