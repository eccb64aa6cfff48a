#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_tune_cls.py > gpurun_out/tune_cls.txt 2>&1; cat gpurun_out/tune_cls.txt
# row-plan kernel: full capture on the weighted per-fragment sample
ncu --set full --clock-control none --import-source on -k regex:"k_alloc_rows" -s 4 -c 1 -o gpurun_out/prof_r02_rows -f python bench.py --weights --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-gates --no-extras > gpurun_out/ncu_rows.log 2>&1
tail -2 gpurun_out/ncu_rows.log
