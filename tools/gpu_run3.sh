#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_linear.py -x -q 2>&1 | tail -15
timeout 200 python tools/linear_time.py
timeout 200 python tools/linear_time.py 4096 4096 4096
