#!/bin/bash
# 2-GPU checks of the class-plan path: parity (NCCL + fused peer-memory exchange), CLI -gpus 2, scaling bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log; tail -4 gpurun_out/pytest_multi.log
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k two_gpu --timeout 300 > gpurun_out/pytest_cli2.log 2>&1; tail -3 gpurun_out/pytest_cli2.log
for L in perfragment collapsed; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --layout $L > gpurun_out/scale2_$L.json 2>gpurun_out/scale2_$L.err
  tail -2 gpurun_out/scale2_$L.err
  python - $L <<'PY'
import json,sys
L=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/scale2_{L}.json").read().strip().split("\n")[-1]); r=d["roofline"]
    print("N 2",L,"value",d["value"],"sweeps/s",round(d["sweeps_per_s"],1),"alloc_ms",round(r["avg_launch_ms"],4),"gamma_ms",round(r["gamma_avg_launch_ms"],4),"step_ms",round(d["ms_per_step"],3),"e2e",d["e2e"] and round(d["e2e"]["sweeps_per_s"],1), r.get("per_rank"))
except Exception as e: print("N 2",L,"failed",e)
PY
done
