#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log; tail -4 gpurun_out/pytest_gpu2.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --no-extras > gpurun_out/bench_2gpu_weak.json 2> gpurun_out/bench_2gpu_weak.err; echo "bench rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --batch 8 --batch-sweeps 4096 > gpurun_out/bench_2gpu_batch.json 2> gpurun_out/bench_2gpu_batch.err; echo "batch rc=$?"; tail -2 gpurun_out/bench_2gpu_batch.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu_weak.json").read().strip().split("\n")[-1]); r=d["roofline"]
print("weak N=2 sweeps/s", round(d["sweeps_per_s"],1), "alloc_ms", r["avg_launch_ms"], "gamma_ms", r["gamma_avg_launch_ms"], "sweep_ms", r["sweep_ms"], "e2e", round(d["e2e"]["sweeps_per_s"],1), "gates", d["gates"]["ok"])
b=json.loads(open("gpurun_out/bench_2gpu_batch.json").read().strip().split("\n")[-1])
print("batch", {k:b[k] for k in ("value","wall_s","sweeps_per_s_per_gpu","gpu_pipeline_s_median","prep_s_median")})
PY
