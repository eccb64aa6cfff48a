#!/bin/bash
# parity of the quad kernel, then A/B timing against the two-rows-per-lane kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_cli.py -x -q -m gpu > gpurun_out/pytest_seg4.log 2>&1; tail -4 gpurun_out/pytest_seg4.log
bash tools/gpu_quick2.sh "MMQ_X=0" "MMQ_SEG_KERNEL=1" "MMQ_SEG_OCC=3" "MMQ_DEBUG_DMAX=6" "MMQ_DEBUG_DMIN=7" "MMQ_DEBUG_DMAX=2" "MMQ_DEBUG_DMIN=13"
for v in "MMQ_X=0" "MMQ_SEG_KERNEL=1"; do
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
ncu --set full --clock-control none --import-source on -k regex:k_alloc_seg4 -s 5 -c 1 -o gpurun_out/prof_seg4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
