#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collapse.py -x -q > gpurun_out/cov_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cov_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cov_gemm -s 1 -c 1 -o gpurun_out/prof_cov -f python tools/gpu_cov_prof.py > gpurun_out/ncu_cov.log 2>&1
tail -3 gpurun_out/ncu_cov.log; tail -5 gpurun_out/cov_pytest.log
