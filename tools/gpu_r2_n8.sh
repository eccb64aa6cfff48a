#!/bin/bash
# 8 GPUs: the default bench line with its gates (weak scaling), config 4 (one weighted 200M-fragment sample split over the ranks),
# config 5 (64 samples, 8 per GPU)
N=${1:-8}
O=gpurun_out
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $T --master-port 29600 bench.py --gpus $N > $O/r02_bench_${N}gpu_weak.json 2> $O/bench_${N}gpu_weak.err; echo "weak rc=$?"; tail -2 $O/bench_${N}gpu_weak.err
if [ "$2" = "c4" ]; then timeout 1200 $T --master-port 29601 bench.py --gpus $N --scaling strong --fragments-total 200000000 --weights --steps 10 > $O/r02_bench_${N}gpu_c4.json 2> $O/bench_${N}gpu_c4.err; echo "c4 rc=$?"; tail -2 $O/bench_${N}gpu_c4.err; fi
timeout 1200 $T --master-port 29602 bench.py --gpus $N --batch $((8 * N)) --batch-per-gpu 4 > $O/r02_bench_${N}gpu_batch.json 2> $O/bench_${N}gpu_batch.err; echo "batch rc=$?"; tail -2 $O/bench_${N}gpu_batch.err
python - <<PY
import json
for nm in ("weak", "c4"):
    try:
        d=json.loads(open("gpurun_out/r02_bench_${N}gpu_%s.json" % nm).read().strip().split("\n")[-1]); r=d["roofline"]
        print(nm, "N=$N sweeps/s", round(d["sweeps_per_s"],1), "value %.4g" % d["value"], "sweep_ms", r["sweep_ms"], "alloc", r["avg_launch_ms"], "gamma", r["gamma_avg_launch_ms"], "e2e", round(d["e2e"]["sweeps_per_s"],1), "gates", d["gates"]["ok"], {k: v.get("ok") for k, v in d["gates"].items() if isinstance(v, dict)})
    except Exception as e: print(nm, "failed", e)
try:
    b=json.loads(open("gpurun_out/r02_bench_${N}gpu_batch.json").read().strip().split("\n")[-1]); print("batch", {k:b[k] for k in ("value","wall_s","sweeps_per_s_per_gpu","gpu_pipeline_s_median","prep_s_median")})
except Exception as e: print("batch failed", e)
PY
