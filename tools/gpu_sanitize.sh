#!/bin/bash
# memcheck of the class-plan and segment kernels on the small parity problems
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "class_plan or by_length or smoke or explicit_class" > gpurun_out/sanitize.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize.log | tail -8
