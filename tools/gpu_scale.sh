#!/bin/bash
# N-GPU scaling bench as the driver launches it: usage gpu_scale.sh N1 N2 ...
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for N in "$@"; do
  if [ $N -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2>gpurun_out/scale_$N.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.json 2>gpurun_out/scale_$N.err; fi
  tail -2 gpurun_out/scale_$N.err
  python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/scale_{N}.json").read().strip().split("\n")[-1]); r=d["roofline"]
    print("N",N,"value",d["value"],"sweeps/s",round(d["sweeps_per_s"],1),"alloc_ms",round(r["avg_launch_ms"],4),"step_ms",round(d["ms_per_step"],3),"e2e",d["e2e"] and round(d["e2e"]["sweeps_per_s"],1), d["config"].get("exchange"))
except Exception as e: print("N",N,"failed",e)
PY
done
N="${@: -1}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --nccl-only --no-e2e 2>gpurun_out/scale_nccl.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N $N nccl-only sweeps/s', round(d['sweeps_per_s'],1), 'step_ms', round(d['ms_per_step'],3))"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
