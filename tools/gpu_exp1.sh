#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --fragments 3000000 --no-cpu-baseline --no-e2e"
$B --transposed > gpurun_out/exp_transposed.json 2>gpurun_out/exp.err
$B --layout perfragment_unsorted > gpurun_out/exp_unsorted.json 2>>gpurun_out/exp.err
$B --layout collapsed > gpurun_out/exp_collapsed.json 2>>gpurun_out/exp.err
$B --weights > gpurun_out/exp_weights.json 2>>gpurun_out/exp.err
ncu --set full --clock-control none --import-source on -k regex:k_alloc -s 20 -c 2 -o gpurun_out/prof_alloc_r1a $B > gpurun_out/ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_r1a.csv $B > /dev/null 2>&1
for f in transposed unsorted collapsed weights; do python - <<PY
import json
d=json.load(open("gpurun_out/exp_$f.json"))
r=d["roofline"]
print("$f", "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", round(r["avg_launch_ms"],4), "gamma_ms", round(r["gamma_avg_launch_ms"],4), "GB/s", round(r["achieved"],1), "m", d["config"]["classes_per_gpu"], "nnz", d["config"]["nnz_per_gpu"])
PY
done
tail -3 gpurun_out/exp.err
