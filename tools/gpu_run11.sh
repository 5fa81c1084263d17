#!/bin/bash
mkdir -p gpurun_out
SIZES="256 256 256 512 512 512 1024 1024 1024 1536 1536 1536 2048 2048 2048 256 4096 4096 64 512 4096 128 1024 8192"
for dt in f32 f64; do
  echo "== $dt split-K auto"; timeout 300 python tools/gemm_small_time.py $dt $SIZES 2>&1 | tail -9
  echo "== $dt split-K off";  WK_GEMM_SPLITK=1 timeout 300 python tools/gemm_small_time.py $dt $SIZES 2>&1 | tail -9
done
