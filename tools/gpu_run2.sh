#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.log 2>&1; echo "bench exit $?" >> gpurun_out/bench3.log
tail -3 gpurun_out/bench3.log | cut -c1-1500
