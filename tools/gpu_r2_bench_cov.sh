#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_cov.json 2> gpurun_out/bench_cov.err; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cov_gemm -s 1 -c 1 -o gpurun_out/prof_cov2 -f python tools/gpu_cov_prof.py > gpurun_out/ncu_cov2.log 2>&1
tail -2 gpurun_out/ncu_cov2.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_cov.json").read().strip().splitlines()[-1])
print("sweeps/s", d["sweeps_per_s"], "e2e", d["e2e"]["sweeps_per_s"], "frac", d["roofline"]["frac"])
print("posterior", json.dumps(d["gates"].get("posterior")))
print("gates ok", d["gates"]["ok"])
print("trace_cov", json.dumps(d.get("trace_cov"))[:1500])
PY
tail -5 gpurun_out/bench_cov.err
