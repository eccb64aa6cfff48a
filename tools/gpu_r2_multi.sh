#!/bin/bash
# multi-GPU: parity tests, then the bench line (with its gates) at N = $1 (default 2): weak scaling on the collapsed C2 shards,
# and config 4 (strong scaling: ONE weighted 200M-fragment sample split over the ranks) when $2 = c4
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log; tail -4 gpurun_out/pytest_multi.log; fi
run() { # name, extra args
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N $2 > gpurun_out/bench_${N}gpu_$1.json 2> gpurun_out/bench_${N}gpu_$1.err; echo "bench $1 N=$N rc=$?"; tail -3 gpurun_out/bench_${N}gpu_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_$1.json").read().strip().split("\n")[-1])
    r=d.get("roofline") or {}
    print("$1 N=$N value %.4g"%d["value"], "sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", r.get("avg_launch_ms"), "gamma_ms", r.get("gamma_avg_launch_ms"), "sweep_ms", r.get("sweep_ms"), "e2e", d["e2e"] and round(d["e2e"].get("sweeps_per_s",0),1), r.get("per_rank"))
    print("  gates", json.dumps(d.get("gates")))
except Exception as e: print("failed", e)
PY
}
run weak ""
if [ "$2" = "c4" ]; then run c4 "--scaling strong --fragments-total 200000000 --weights --steps 10"; fi
