#!/bin/bash
# round 2, first contact: GPU tests, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
for f in ("ours","reference"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().split("\n")[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g"%d["value"], "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", r.get("avg_launch_ms"), "gamma_ms", r.get("gamma_avg_launch_ms"), "gap_us", r.get("launch_gap_us_per_sweep"), "frac", r.get("frac"), "e2e", d["e2e"] and round(d["e2e"].get("sweeps_per_s",0),1), "cpu", d.get("cpu_baseline",{}).get("sweeps_per_s"))
        print("  gates", json.dumps(d.get("gates")))
        print("  weighted", json.dumps(d.get("perfragment_weighted")))
    except Exception as e: print(f,"failed",e)
PY
