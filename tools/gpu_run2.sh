#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/tc_check.py f32 > gpurun_out/tc_check_f32.log 2>&1; echo "exit $?" >> gpurun_out/tc_check_f32.log
grep -v "^OK" gpurun_out/tc_check_f32.log | tail -30
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
