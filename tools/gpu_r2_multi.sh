#!/bin/bash
# multi-GPU: parity tests on 2 GPUs, then the bench line (with its gates) at N = 2 [and N = $1 if given]
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log; tail -15 gpurun_out/pytest_multi.log
for N in 2 $1; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N $BENCH_ARGS > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N rc=$?"; tail -5 gpurun_out/bench_${N}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().split("\n")[-1])
    r=d.get("roofline") or {}
    print("N=$N value %.4g"%d["value"], "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", r.get("avg_launch_ms"), "gamma_ms", r.get("gamma_avg_launch_ms"), "gap_us", r.get("launch_gap_us_per_sweep"), "e2e", d["e2e"] and round(d["e2e"].get("sweeps_per_s",0),1), r.get("per_rank"))
    print("  gates", json.dumps(d.get("gates")))
except Exception as e: print("failed", e)
PY
done
