#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --layout collapsed > gpurun_out/exp_collapsed_full.json 2>gpurun_out/exp.err; tail -2 gpurun_out/exp.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/exp_collapsed_full.json")); r=d["roofline"]
print("collapsed full: sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "m", d["config"]["classes_per_gpu"], "nnz", d["config"]["nnz_per_gpu"])
PY
