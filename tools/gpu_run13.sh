#!/bin/bash
mkdir -p gpurun_out
for dt in c64 c128; do timeout 300 python tools/gemm_time.py $dt 2048 4096 8192 2>&1 | tail -3; done
for cfg in "2 16" "4 16" "4 8" "8 16" "2 32"; do set -- $cfg
  echo "== e2e NJ=$1 panels=$2"; WK_E2E_NJ=$1 WK_E2E_PANELS=$2 timeout 600 python bench.py --steps 3 --warmup 3 --quick --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
