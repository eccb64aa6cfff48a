#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for mode in 2 1; do
MMQ_P2P_MODE=$mode MMQ_P2P_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/trace_$N.json 2>gpurun_out/trace_$N.err
echo "mode $mode"; grep "p2p trace" gpurun_out/trace_$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f"gpurun_out/trace_{N}.json").read().strip().split("\n")[-1]); r=d["roofline"]
print("N",N,"sweeps/s",round(d["sweeps_per_s"],1),"step_ms",round(d["ms_per_step"],3), r.get("per_rank"))
PY
done
MMQ_P2P_MODE=2 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
