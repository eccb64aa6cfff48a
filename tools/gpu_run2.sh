#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --fragments 3000000 --no-cpu-baseline --no-e2e"
$B > gpurun_out/exp_default.json 2>gpurun_out/exp.err
$B --layout perfragment_byclass > gpurun_out/exp_unsorted.json 2>>gpurun_out/exp.err
$B --weights > gpurun_out/exp_weights.json 2>>gpurun_out/exp.err
python bench.py --no-cpu-baseline > gpurun_out/exp_full.json 2>>gpurun_out/exp.err
ncu --set full --clock-control none --import-source on -k regex:k_alloc -s 20 -c 1 -o gpurun_out/prof_alloc_r1f $B > gpurun_out/ncu.log 2>&1
for f in default unsorted weights full; do python - <<PY
import json
d=json.load(open("gpurun_out/exp_$f.json"))
r=d["roofline"]
print("$f", "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "GB/s", round(r["achieved"],1), "frac", round(r["frac"],3), "m", d["config"]["classes_per_gpu"], "nnz", d["config"]["nnz_per_gpu"], "e2e", d["e2e"] and round(d["e2e"]["sweeps_per_s"],1))
PY
done
tail -3 gpurun_out/exp.err
