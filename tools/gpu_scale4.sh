#!/bin/bash
# 4-GPU weak-scaling check of the bench (fused peer-memory exchange and NCCL)
mkdir -p gpurun_out
N=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale4.json 2>gpurun_out/scale4.err; echo "rc=$?"; tail -2 gpurun_out/scale4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --nccl-only --no-e2e > gpurun_out/scale4_nccl.json 2>gpurun_out/scale4_nccl.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale4_ref.json 2>gpurun_out/scale4_ref.err; echo "ref rc=$?"; tail -c 300 gpurun_out/scale4_ref.json
python - <<'PY'
import json
for f in ("scale4","scale4_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1]); r=d["roofline"]
        print(f,"N",d["n_gpus"],"value %.4g"%d["value"],"sweeps/s",round(d["sweeps_per_s"],1),"alloc_ms",round(r["avg_launch_ms"],4),"gamma_ms",round(r["gamma_avg_launch_ms"],4),"step_ms",round(d["ms_per_step"],3),"e2e",d["e2e"] and round(d["e2e"]["sweeps_per_s"],1), r.get("per_rank"))
    except Exception as e: print(f,"failed",e)
PY
