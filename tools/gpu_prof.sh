#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_alloc_seg4 -s 5 -c 1 -o gpurun_out/prof_seg4 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
