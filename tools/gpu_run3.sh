#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
