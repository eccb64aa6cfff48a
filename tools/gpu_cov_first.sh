#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_cov_first.py > gpurun_out/cov_first.log 2>&1; echo "cov_first rc=$?" >> gpurun_out/cov_first.log
timeout 600 python -m pytest tests/test_gpu_collapse.py -x -q > gpurun_out/cov_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cov_pytest.log
tail -30 gpurun_out/cov_first.log; tail -15 gpurun_out/cov_pytest.log
