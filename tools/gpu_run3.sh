#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/tc_check.py f64 > gpurun_out/tc_check_f64.log 2>&1; echo "exit $?" >> gpurun_out/tc_check_f64.log
grep -v "^OK" gpurun_out/tc_check_f64.log | tail -14
