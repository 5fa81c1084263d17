#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_xor.py -x -q 2>&1 | tail -5
timeout 120 python tools/xor_time.py
