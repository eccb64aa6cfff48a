#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
bash tools/gpu_quick.sh "MMQ_SEG_KERNEL=2" "MMQ_SEG_KERNEL=1"
B="python bench.py --steps 5 --warmup 3 --fragments 3000000 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:k_alloc -s 20 -c 1 -o gpurun_out/prof_alloc_r1g $B > gpurun_out/ncu.log 2>&1
