#!/bin/bash
# 2-GPU checks: NCCL path parity + scaling bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log; tail -5 gpurun_out/pytest_multi.log
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k two_gpu > gpurun_out/pytest_cli2.log 2>&1; tail -3 gpurun_out/pytest_cli2.log
for N in 1 2; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2>gpurun_out/scale_$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.json 2>gpurun_out/scale_$N.err; fi
  tail -2 gpurun_out/scale_$N.err
  python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/scale_{N}.json").read().strip().split("\n")[-1]); r=d["roofline"]
    print("N",N,"value",d["value"],"sweeps/s",round(d["sweeps_per_s"],1),"alloc_ms",round(r["avg_launch_ms"],4),"step_ms",round(d["ms_per_step"],3),"e2e",d["e2e"] and round(d["e2e"]["sweeps_per_s"],1))
except Exception as e: print("N",N,"failed",e)
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --nccl-only --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N 2 nccl-only sweeps/s', round(d['sweeps_per_s'],1), 'step_ms', round(d['ms_per_step'],3))"
