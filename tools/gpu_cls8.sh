#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_cli.py -x -q -m gpu --timeout 300 > gpurun_out/cls_parity.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/cls_parity.log
MMQ_CREATE_TIMING=1 python tools/time_create2.py 2>&1 | tail -24
python bench.py --layout collapsed --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q.json 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
print("collapsed | sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "e2e", d["e2e"])
PY
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q.json 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/q.json")); r=d["roofline"]
print("perfragment | sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "e2e", d["e2e"])
PY
