#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --layout collapsed --transcripts 1000 --fragments 100000 > gpurun_out/exp_c1.json 2>gpurun_out/exp.err; tail -2 gpurun_out/exp.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/exp_c1.json")); r=d["roofline"]
print("C1 collapsed: sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "step_ms", round(d["ms_per_step"],4), "m", d["config"]["classes_per_gpu"], "nnz", d["config"]["nnz_per_gpu"])
PY
