#!/bin/bash
# 8 GPUs: default bench (fused gather headline, shared-upload e2e, axpy per rank)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_g8.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_g8.log | cut -c1-1500
