#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_q2.json 2> $O/bench_q2.err; echo "ours rc=$?"; tail -2 $O/bench_q2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_q2.json").read().strip().splitlines()[-1])
print("sweeps/s", d["sweeps_per_s"], "e2e", d["e2e"]["sweeps_per_s"], "frac", d["roofline"]["frac"])
print("posterior", json.dumps(d["gates"].get("posterior")))
print("gates ok", d["gates"]["ok"], "cpu", d["cpu_baseline"]["sweeps_per_s"])
print("trace_cov", d["trace_cov"]["ms"], d["trace_cov"]["roofline"]["frac"], d["trace_cov"]["gate"])
PY
