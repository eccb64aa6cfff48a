#!/bin/bash
# GPU tests + a short bench line (no CPU arm) + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().split("\n")[-1])
r=d.get("roofline") or {}
print("value %.4g"%d["value"], "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", r.get("avg_launch_ms"), "gamma_ms", r.get("gamma_avg_launch_ms"), "gap_us", r.get("launch_gap_us_per_sweep"), "frac", r.get("frac"), "e2e", d["e2e"] and round(d["e2e"].get("sweeps_per_s",0),1))
print("  gates", json.dumps(d.get("gates")))
print("  weighted", json.dumps(d.get("perfragment_weighted")))
PY
bash tools/gpu_r2_prof.sh
