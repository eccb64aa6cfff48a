#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --layout collapsed > gpurun_out/exp_collapsed_full.json 2>gpurun_out/exp.err; tail -2 gpurun_out/exp.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/exp_collapsed_full.json")); r=d["roofline"]
print("collapsed full: sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "m", d["config"]["classes_per_gpu"], "nnz", d["config"]["nnz_per_gpu"])
PY
ncu --set full --clock-control none --import-source on -k regex:k_alloc -s 10 -c 1 -o gpurun_out/prof_alloc_collapsed_r1b python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --layout collapsed --fragments 10000000 > gpurun_out/ncu.log 2>&1
